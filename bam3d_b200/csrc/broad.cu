// Broad-phase kernels (sweep and prune, BVH queries) — see broad.h.
#include "broad.h"
