// oracle/shim/itpp/base/smat.h -- TEST INFRASTRUCTURE ONLY: see vec.h
#pragma once
#include "vec.h"
