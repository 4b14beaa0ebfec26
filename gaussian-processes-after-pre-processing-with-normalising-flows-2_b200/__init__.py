"""flowk - B200-native (sm_100a) kernels and nn.Module front-end for the mAR-SCF flow-step path.

The directory name carries the reference's name; `import flowk` (see flowk.py at the repo
root) registers this package under the importable alias `flowk`.

Layout mirrors the reference's module paths for the hot path only:
    flowk.flow_modules.common_modules   squeeze2d, unsqueeze2d, SqueezeLayer, InvertibleConv1x1,
                                        Actnormlayer, Split2dMsC, TupleFlip, GaussianDiag
    flowk.flow_modules.affine_coupling  Conv2dZeros, Conv2d, NN_net, AffineCoupling
    flowk.flow_modules.mixlogcdf_coupling  MixLogCDFCoupling
    flowk.flow_modules.mixlogcdf_nn     WNConv2d, NN, ConvAttnBlock, GatedAttn, GatedConv, Rescale
    flowk.flow_modules.log_dist         mixture_log_pdf / mixture_log_cdf / mixture_inv_cdf / inverse
    flowk.flow_modules.misc             cpd_sum, cpd_mean
    flowk.marscf                        FlowStep, FlowNet, MarScfFlow
    flowk.ops                           torch.library custom ops over the C ABI (include/flowk.h)
    flowk.sharding                      batch-sharded evaluation / training helpers (torch.distributed)
"""
import os as _os

import torch as _torch

from . import _lib  # noqa: F401  fails loudly when libflowk.so has not been built

# Parity budget is 1e-4 relative in fp32 (BASELINE.json north_star): TF32 tensor-core convolutions
# (torch's default for cuDNN) miss it by ~10x, so the library conditioner convs run in full fp32 unless
# the user opts back in.
if _os.environ.get("FLOWK_ALLOW_TF32", "0") != "1":
    _torch.backends.cudnn.allow_tf32 = False
    _torch.backends.cuda.matmul.allow_tf32 = False
from . import ops  # noqa: F401

__all__ = ["ops"]
