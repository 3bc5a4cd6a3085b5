// placeholder until the tiled push lands
#include "common.cuh"
namespace ib200 {
int try_push_tiled(int, const KParams &, int, const void *, const void *, void *, cudaStream_t) { return 0; }
}
