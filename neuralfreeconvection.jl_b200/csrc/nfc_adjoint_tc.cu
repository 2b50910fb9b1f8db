// nfc_adjoint_tc.cu -- separate translation unit of the tensor-core backward kernel (nfc_adjoint_tc.cuh) and its launcher
#include "nfc_adjoint_tc.cuh"

namespace nfc {

int adjoint_tc_setup(nfc_handle* h) {
    NFC_CUDA(h, cudaFuncSetAttribute(atc::adjoint_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, atc::ATC_SMEM));
    return 0;
}

int adjoint_tc_launch(nfc_handle* h, const AdjParams& A, const atc::AtcScales& sc, int64_t nb, cudaStream_t st) {
    const int grid = (int)std::min<int64_t>(h->num_sms, (nb + tc::TILE - 1) / tc::TILE);
    atc::adjoint_tc_kernel<<<grid, tc::NTHREADS, atc::ATC_SMEM, st>>>(h->d_wimg.p, A, sc);
    NFC_CUDA(h, cudaGetLastError());
    return 0;
}

}  // namespace nfc
