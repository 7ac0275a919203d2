// ks_bitpar.cu — placeholder until the bit-parallel kernels land (next commit).
#include "ks_bitpar.h"
namespace ks {
cudaError_t bitpar_build(BitparIndex& ix, const NodeTable& nt, const int64_t*, cudaStream_t) {
    ix.N = nt.N; ix.Npad = nt.Npad; ix.W = nt.W; ix.valid = false;
    return cudaSuccess;
}
bool bitpar_profitable(const BitparIndex&, uint32_t) { return false; }
cudaError_t bitpar_select(BitparIndex&, const SelectLaunch&, const int64_t*, cudaEvent_t) { return cudaErrorNotSupported; }
void bitpar_release(BitparIndex& ix) { if (ix.blob) cudaFree(ix.blob); if (ix.order) cudaFree(ix.order); ix = BitparIndex(); }
}
