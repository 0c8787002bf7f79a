#pragma once
#include "csr_matrix.h"
