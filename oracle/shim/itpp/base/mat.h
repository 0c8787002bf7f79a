// oracle/shim/itpp/base/mat.h -- TEST INFRASTRUCTURE ONLY: see vec.h
#pragma once
#include "vec.h"
