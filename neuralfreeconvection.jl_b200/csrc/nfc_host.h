// nfc_host.h -- host-side state of one nfc_handle (device buffers, stream, events) shared by nfc_api.cu and nfc_adjoint.cuh
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/nfc.h"

namespace nfc {

enum Arch { ARCH_DEFAULT = 0, ARCH_WIDER = 1, ARCH_DEEPER = 2, ARCH_CONV2 = 3, ARCH_CONV4 = 4, ARCH_NONE = -1 };

template <class T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(&p, n * sizeof(T));
        if (e == cudaSuccess) cap = n;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

namespace atc {
// operand scales of the tensor-core backward kernel (nfc_adjoint_tc.cuh): exact powers of two chosen in nfc_set_weights
struct AtcScales {
    float c_g2, c_g1;                 // accumulator -> next A operand of the transposed chain
    float inv_u;                      // ubar accumulator -> ubar (times 1/sc of the column)
    float s_d3;                       // scale of the output-layer delta operand (2^14)
    float inv_sd2, inv_sd1;           // operand scale -> true delta (times 1/sc of the column)
    float inv_sa0, inv_sa1, inv_sa2;  // the staged activations carry the forward operand scales: undone when the accumulators are flushed
};
}  // namespace atc

struct AdjParams;
}  // namespace nfc

struct nfc_handle;
namespace nfc {
// [grad; loss sums] all-reduce over the handle's communicator + normalisation by the global residual count (nfc_api.cu)
int comm_allreduce_grad(nfc_handle* h, float* grad_dev, double* loss_sum_dev, cudaStream_t st);
int adjoint_tc_setup(nfc_handle* h);
int adjoint_tc_launch(nfc_handle* h, const AdjParams& A, const atc::AtcScales& sc, int64_t nb, cudaStream_t st);
}  // namespace nfc

struct nfc_handle {
    nfc_config cfg;
    std::vector<float> saveat;
    nfc::Arch arch = nfc::ARCH_NONE;
    int device = 0, num_sms = 0;
    int64_t n_params = 0, n_packed = 0;
    // Nz < 32 (--grid-points): the caller's Chain and columns are embedded in the 32-level kernels (zero-padded weights, levels [0, nz));
    // cfg holds the PADDED Chain, theta_map[user index] = index in the padded parameter vector (nfc_api.cu)
    int nz = 32;
    int64_t n_params_user = 0;
    std::vector<int64_t> theta_map;
    bool in_pad = false;                    // a host-buffer call is running on padded copies of the caller's arrays
    bool weights_set = false;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    cudaStream_t heavy_stream = nullptr;    // high priority: the longest columns of a loss_grad chunk, one column per CTA (nfc_adjoint.cuh)
    cudaEvent_t ev_heavy[2] = {nullptr, nullptr};
    int64_t adj_heavy = 48;                 // how many (0 = none); needs the previous call's per-column step counts
    cudaEvent_t ev_chunk[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    bool ev_fwd_valid = false, ev_adj_valid = false;
    nfc::DevBuf<float> d_theta, d_wpack, d_saveat, d_k1, d_dt;
    nfc::DevBuf<float> e_theta, e_wpack;
    nfc::DevBuf<float> d_rkctab;           // RKC coefficient table (alg = NFC_ALG_RKC2)
    nfc::DevBuf<int> s_nfe;
    int rkc_mmax = 2;   // ensemble members (nfc_solve_ensemble)
    nfc::DevBuf<unsigned char> d_wimg;      // tensor-core weight image (default Chain only)
    bool use_tc = false;
    bool tc_record = true;                  // forward (RECORD) pass of loss_grad on the tensor-core kernels (FP16x2 split)
    int tc_tiles = 0;                       // FP16x2 solve: 0 = pick by batch size, 1 = one tile per CTA (nfc_tc.cuh), 2 = two tiles in flight (nfc_tc2.cuh)
    int64_t tc2_min_columns = 98304;        // auto: two tiles from this batch size on
    int e2e_pieces = 0;                     // nfc_solve: pieces whose device->host copies overlap the next piece's solve (0 = pick by batch size)
    double tape_gib = 0.0;                  // scratch budget of nfc_loss_grad's tape in GiB (0 = default 48)
    bool adj_tc = true;                     // batch backward pass of loss_grad on the tensor cores (default Chain, FP16x2; NFC_ADJ_TC=0: FP32 SIMT kernel)
    nfc::atc::AtcScales atc_scales{};
    float fp16_scales[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};   // tc::Fp16Scales of the current weights (nfc_set_weights), for re-packing the image
    bool range_fallback = true;             // host-buffer calls re-solve ST_RANGE columns on a split / kernel without the FP16 range bound
    int64_t range_resolved = 0;             // columns the last host-buffer call re-solved that way
    bool tc_fp16 = false;                   // tensor-core split: false = BF16x3 (6 MMAs / product), true = FP16x2 (3 MMAs, scaled operands)
    // longest-first scheduling: per-column attempted steps of the previous forward call with the same batch size
    nfc::DevBuf<int> d_work, d_order, d_hist;
    int64_t sched_B = -1;       // batch size d_work is valid for (-1: none)
    std::vector<bool> sched_valid = std::vector<bool>(8, false);   // per piece of a host-level batch: history present
    bool use_sched = true;
    bool record_sched = false;     // set by nfc_loss_grad around a RECORD pass that covers the whole batch: it may use and feed the step history
    nfc::DevBuf<unsigned long long> d_counters;
    // host-call staging
    nfc::DevBuf<float> s_T0, s_colp, s_out, s_true, s_grad;
    nfc::DevBuf<int> s_nacc, s_nrej, s_status;
    nfc::DevBuf<double> s_loss;
    // adjoint scratch
    nfc::DevBuf<float> a_resid, a_tape, a_gacc;
    nfc::DevBuf<int> a_tlen, a_adjsteps, a_count, a_flag;
    int64_t a_count_B = -1;                    // compact tape: a_count holds the accepted steps per column of the last call on this batch size
    bool a_count_valid = false;
    nfc::DevBuf<long long> a_off;              // compact tape: first row of every column, per chunk (nfc_adjoint.cuh)
    std::vector<long long> a_off_host;         // the last chunk's offsets (nfc_get_tape); empty: dense tape
    int64_t a_tape_rows = 0;                   // rows of the tape buffer the last call used
    int tape_mode = 0;                         // 0 = compact tape when the dense one exceeds the budget, 1 = dense, 2 = compact
    bool count_pass = false;                   // launch_solve: pick the kernels of the RECORD pass for a non-RECORD solve
    int64_t a_adjsteps_n = 0;      // a_adjsteps: attempted backward steps per column of the last nfc_loss_grad batch (all chunks)
    int64_t fwd_small_max = 592;   // (4 per SM: tools/gpu_fwd_crossover.py) forward solves / RECORD passes of at most this many columns (default Chain) run one column per CTA (nfc_forward_small.cuh)
    int64_t adj_small_max = 2000;  // loss_grad chunks of at most this many columns (default Chain) use the one-column-per-CTA backward kernel (half of it once the step history is known)
    int64_t a_adjhist_B = -1;      // batch size that history belongs to; it is the next call's longest-first key
    bool a_adjhist_valid = false;
    nfc::DevBuf<double> a_gsum;
    // multi-GPU: NCCL communicator attached by nfc_comm_init (null: single GPU); the all-reduce of nfc_loss_grad runs on it
    void* nccl_comm = nullptr;
    int comm_rank = 0, comm_world = 1;
    float last_allreduce_ms = 0.f;
    cudaEvent_t ev_ar[2] = {nullptr, nullptr};
    int64_t last_columns = 0, last_launches = 0;
    std::string err;
};

namespace nfc {

inline std::string& create_error() { static thread_local std::string e; return e; }

inline int fail(nfc_handle* h, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf; else create_error() = buf;
    return code;
}

#define NFC_CUDA(h, call)                                                                              \
    do {                                                                                               \
        cudaError_t e__ = (call);                                                                      \
        if (e__ != cudaSuccess) return nfc::fail(h, -10, "CUDA error %s at %s:%d (%s)", cudaGetErrorName(e__), __FILE__, __LINE__, cudaGetErrorString(e__)); \
    } while (0)

}  // namespace nfc
