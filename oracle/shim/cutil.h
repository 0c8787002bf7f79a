// oracle/shim/cutil.h -- TEST INFRASTRUCTURE ONLY.  Empty stand-in for the CUDA SDK 3.x header src_thermal/main2.cu:20 includes
// (nothing of it is used there); lets the unchanged driver be compiled against the shim in tests/test_shim.py.
#pragma once
