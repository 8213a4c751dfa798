// mpcqp_ric_consts.h -- launch / workspace constants of the stage-wise solver shared by its kernels and the host side.
#pragma once
namespace mpcqp {
constexpr int RIC_GAIN = 84;        // workspace doubles per stage: 6 impulse components x 14 (13 coefficients + pad)
constexpr int RIC_ADM = 18;         // workspace doubles of interior-point state per foot-step (struct of arrays): f (3), y (6), step df (3), dy (6)
#ifndef MPCQP_RIC_WARPS
#define MPCQP_RIC_WARPS 1
#endif
constexpr int RIC_WARPS = MPCQP_RIC_WARPS;      // warps per CTA
constexpr int RIC_PER_CTA = 2 * RIC_WARPS;      // robots per CTA (half a warp each)
}  // namespace mpcqp
