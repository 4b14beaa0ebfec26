"""Convolutional LSTM over the channel sequence, with the reference's module tree / state-dict keys
(mar_prior/lstm.py:7-43, mar_prior/convolutional_rnn/module.py:13-218,378-404, functional.py:30-52,98-160,248-275).

Plain torch (the prior is outside the flow-step hot path): stacked layers, each run over the
whole sequence; gates = conv_same(x_t, W_ih) + conv_same(h_{t-1}, W_hh), chunked (i, f, g, o).  The input-to-hidden
convolutions of a layer are batched over time (they do not depend on the recurrence)."""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


class Conv2dLSTM(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, num_layers=1, bias=True, batch_first=False,
                 dropout=0., bidirectional=False, stride=1, dilation=1, groups=1):
        super().__init__()
        assert not bidirectional and stride == 1 and groups == 1 and bias, "only the configuration the prior uses"
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.dilation = int(kernel_size), int(dilation)
        self.num_layers, self.batch_first, self.dropout = num_layers, batch_first, dropout
        k = self.kernel_size
        for layer in range(num_layers):
            cin = in_channels if layer == 0 else out_channels
            setattr(self, "weight_ih_l%d" % layer, nn.Parameter(torch.empty(4 * out_channels, cin, k, k)))
            setattr(self, "weight_hh_l%d" % layer, nn.Parameter(torch.empty(4 * out_channels, out_channels, k, k)))
            setattr(self, "bias_ih_l%d" % layer, nn.Parameter(torch.empty(4 * out_channels)))
            setattr(self, "bias_hh_l%d" % layer, nn.Parameter(torch.empty(4 * out_channels)))
        stdv = 1.0 / math.sqrt(out_channels)                        # module.py:94-97
        for w in self.parameters():
            w.data.uniform_(-stdv, stdv)

    def _conv(self, x, w, b):
        """'same' zero padding for stride 1 (functional.py:248-272): d*(k-1)/2 on every side for odd k."""
        return F.conv2d(x, w, b, padding=self.dilation * (self.kernel_size - 1) // 2, dilation=self.dilation)

    def forward(self, input, hx=None):
        x = input if self.batch_first else input.transpose(0, 1)     # [B, T, C, H, W]
        B, T = x.shape[0], x.shape[1]
        if hx is None:
            zeros = x.new_zeros(self.num_layers, B, self.out_channels, x.shape[3], x.shape[4])
            hx = (zeros, zeros)
        h_n, c_n = [], []
        for layer in range(self.num_layers):
            w_ih, w_hh = getattr(self, "weight_ih_l%d" % layer), getattr(self, "weight_hh_l%d" % layer)
            b_ih, b_hh = getattr(self, "bias_ih_l%d" % layer), getattr(self, "bias_hh_l%d" % layer)
            gi = self._conv(x.reshape(B * T, *x.shape[2:]), w_ih, b_ih).view(B, T, 4 * self.out_channels, *x.shape[3:])
            h, c = hx[0][layer], hx[1][layer]
            outs = []
            for t in range(T):
                gates = gi[:, t] + self._conv(h, w_hh, b_hh)
                i, f, g, o = gates.chunk(4, 1)
                c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
                h = torch.sigmoid(o) * torch.tanh(c)
                outs.append(h)
            x = torch.stack(outs, dim=1)
            if self.dropout and layer < self.num_layers - 1:
                x = F.dropout(x, self.dropout, self.training)
            h_n.append(h)
            c_n.append(c)
        out = x if self.batch_first else x.transpose(0, 1)
        return out, (torch.stack(h_n, 0), torch.stack(c_n, 0))


class ConvSeqEncoder(nn.Module):
    """conv_embed -> Conv2dLSTM -> conv_out1, every conv applied per time step (mar_prior/lstm.py:7-43)."""

    def __init__(self, input_ch, out_ch, embed_ch, kernel_size=5, dilation=1, num_layers=1, bidirectional=False,
                 dropout=0.0):
        super().__init__()
        self.lstm = Conv2dLSTM(embed_ch, embed_ch, kernel_size, num_layers=num_layers, bidirectional=bidirectional,
                               dilation=dilation, stride=1, dropout=0.0, batch_first=True)
        self.conv_embed = nn.Conv2d(input_ch, embed_ch, kernel_size, stride=1, padding=(1 if kernel_size == 3 else 2))
        self.conv_out1 = nn.Conv2d(embed_ch, out_ch, 3, stride=1, padding=1)
        self.embed_ch, self.out_ch, self.dropout = embed_ch, out_ch, dropout
        self.conv_dropout = nn.Dropout2d(dropout)

    def td_conv(self, x, conv_fn, out_ch):
        b, t = x.size(0), x.size(1)
        y = conv_fn(x.reshape(b * t, x.size(2), x.size(3), x.size(4)))
        if self.dropout > 0:
            y = self.conv_dropout(y)
        return y.view(b, t, out_ch, y.size(2), y.size(3))

    def forward(self, x, lengths=None, hidden=None):
        x2 = self.td_conv(x, self.conv_embed, self.embed_ch)
        outputs, hidden = self.lstm(x2, hidden)
        return self.td_conv(outputs, self.conv_out1, self.out_ch), hidden
