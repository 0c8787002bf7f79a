// oracle/shim/cutil_inline.h -- TEST INFRASTRUCTURE ONLY.  Empty stand-in, see cutil.h (src_thermal/main2.cu:21).
#pragma once
