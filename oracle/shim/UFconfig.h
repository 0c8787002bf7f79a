// oracle/shim/UFconfig.h -- TEST INFRASTRUCTURE ONLY: stand-in, UF_long comes from cs.h
#pragma once
#include "cs.h"
