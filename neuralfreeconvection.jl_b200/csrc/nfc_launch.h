// nfc_launch.h -- the five reference Chains (train_free_convection_nde.jl:125-153) and their launch configurations
#pragma once
#include "nfc_device.cuh"
#include "nfc_host.h"

namespace nfc {

using NetDefault = Net<128, 2, 0>;   // dense-default  Dense(32,128,relu) Dense(128,128,relu) Dense(128,31)
using NetWider = Net<256, 2, 0>;     // dense-wider
using NetDeeper = Net<128, 3, 0>;    // dense-deeper
using NetConv2 = Net<128, 2, 2>;     // conv-2
using NetConv4 = Net<128, 2, 4>;     // conv-4

// weights in shared memory (WS) when they fit next to NW warps of activation scratch, else L1/L2-resident global
template <class N> struct Launch;
template <> struct Launch<NetDefault> { static constexpr bool WS = true;  static constexpr int NW = 8; };
template <> struct Launch<NetWider>   { static constexpr bool WS = false; static constexpr int NW = 8; };   // 330 KB of weights
template <> struct Launch<NetDeeper>  { static constexpr bool WS = true;  static constexpr int NW = 5; };   // 172 KB of weights + 5 x 10 KB
template <> struct Launch<NetConv2>   { static constexpr bool WS = true;  static constexpr int NW = 8; };
template <> struct Launch<NetConv4>   { static constexpr bool WS = true;  static constexpr int NW = 8; };

template <class N> constexpr size_t smem_bytes(int nw) { return sizeof(float) * ((Launch<N>::WS ? N::P_TOTAL : 0) + (size_t)nw * N::S_TOTAL); }

template <class N> int adjoint_setup(nfc_handle* h);
template <class N, bool WS, int NW>
int adjoint_loss_grad(nfc_handle* h, const float* T0, const float* colp, float pref, const float* Ttrue, double* loss_sum_dev, float* grad_dev,
                      int* status, int64_t B, int64_t n_total_columns, cudaStream_t st);

}  // namespace nfc
