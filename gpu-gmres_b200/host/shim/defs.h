// defs.h (B200 shim) -- constants of the reference's src/defs.h:11-47.
#ifndef DEFS_H_
#define DEFS_H_
const int restart = 32, max_iter = 60000;
const float tolerance = 1e-6;
enum PreconditionerType { NONE, DIAG, ILU0, ILUK, AINV };
const float eps = 1e-9;
#endif
