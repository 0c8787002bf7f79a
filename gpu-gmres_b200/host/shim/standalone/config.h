// config.h (B200 shim, STANDALONE fallback) -- launch-shape macros of the reference's src/config.h the drivers use to
// size the SpMV<<<grid, block>>> launch (dim3 grid(1, ceil(n / (BLOCKSIZE/HALFWARP))), block(1, NUMTHREADS)).
#ifndef __CONFIG_H__
#define __CONFIG_H__
#define NUMTHREADS 512
#define BLOCKSIZE 512
#define HALFWARP 16
#define WARPSIZE 32
#define INSPECT 0
#define INSPECT_INPUT 0
#define PADDED_CSR 0
#define CACHE 0
#define TIMER 1
#define SUB_TIMER 0
#endif
