// oracle/shim/itpp/base/math/elem_math.h -- TEST INFRASTRUCTURE ONLY: see ../vec.h
#pragma once
#include "../vec.h"
