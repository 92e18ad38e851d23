#ifndef CQWEN3TTSCUDA_SHIM_H
#define CQWEN3TTSCUDA_SHIM_H
/* Point the header search path at <this repo>/include (unsafeFlags(["-I", ...]) or a copied header). */
#include "qwen3tts_cuda.h"
#endif
