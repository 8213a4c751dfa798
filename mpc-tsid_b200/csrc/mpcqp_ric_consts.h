// mpcqp_ric_consts.h -- launch / workspace constants of the stage-wise solver shared by its kernels and the host side.
#pragma once
#include <cuda_runtime.h>
namespace mpcqp {
constexpr int RIC_GAIN = 84;        // workspace doubles per stage: 6 impulse components x 14 (13 coefficients + pad)
constexpr int RIC_ADM = 18;         // workspace doubles of interior-point state per foot-step (struct of arrays): f (3), y (6), step df (3), dy (6)
#ifndef MPCQP_RIC_WARPS
#define MPCQP_RIC_WARPS 1
#endif
// Debug build (-DMPCQP_CANARY, `make canary`): guard words between the shared-memory arrays of a robot and behind the two
// halves of its workspace slot, checked after every robot (compute-sanitizer is closed on the B200 pool).
#ifdef MPCQP_CANARY
constexpr int RIC_WS_PAD = 2;
#else
constexpr int RIC_WS_PAD = 0;
#endif
__host__ __device__ constexpr int ric_ws_slot_doubles(int cap) { return (RIC_GAIN + 4 * RIC_ADM) * cap + 2 * RIC_WS_PAD; }      // per resident half-warp
constexpr int RIC_WARPS = MPCQP_RIC_WARPS;      // warps per CTA
constexpr int RIC_PER_CTA = 2 * RIC_WARPS;      // robots per CTA (half a warp each)
}  // namespace mpcqp
