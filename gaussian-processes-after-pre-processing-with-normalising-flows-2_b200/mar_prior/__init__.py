from .corr_prior import ChannelPriorMultiScale, ChannelPriorUniScale  # noqa: F401
from .lstm import ConvSeqEncoder, Conv2dLSTM  # noqa: F401
