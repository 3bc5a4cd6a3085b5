// placeholder until the tiled fast paths land
#include "common.cuh"
namespace ib200 {
int try_pull_tiled(const KParams &, int, const void *, const void *, void *, cudaStream_t) { return 0; }
int try_push_tiled(int, const KParams &, int, const void *, const void *, void *, cudaStream_t) { return 0; }
}
