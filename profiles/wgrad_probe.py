"""tcgen05 split-K weight gradient vs fp64 and vs cuDNN / cuBLAS timing."""
import sys, time, torch
sys.path.insert(0, '/root/repo')
import flowk
from flowk import tc_autograd as ta, _lib
dev = torch.device('cuda:0')
torch.manual_seed(0)
def check(b, cin, n, k, h, w, time_it=True):
    x = torch.randn(b, cin, h, w, device=dev); gy = torch.randn(b, n, h, w, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    import ctypes
    tr = ctypes.c_int(0)
    splits = _lib.lib.flowk_conv_wgrad_splits(b, h, w, cin, n, k * k, ctypes.byref(tr))
    tr = bool(tr.value)
    if splits <= 0:
        print("unsupported", (b, cin, n, k, h, w)); return
    part = torch.full((splits, k * k, cin, n) if tr else (splits, k * k, n, cin), float('nan'), device=dev)
    xl = torch.empty_like(x); xr = torch.empty_like(x)
    st = torch.cuda.current_stream().cuda_stream
    def mine_fn(status_ptr=None):
        if k == 3:
            _lib.call("flowk_shift_columns", x.data_ptr(), xl.data_ptr(), xr.data_ptr(), x.numel(), w, st)
        _lib.call("flowk_conv_wgrad", x.data_ptr(), xl.data_ptr(), xr.data_ptr(), gy.data_ptr(), part.data_ptr(), status_ptr,
                  b, h, w, cin, n, k * k, st)
    mine_fn(status.data_ptr())
    torch.cuda.synchronize()
    got = (part.sum(0).permute(2, 1, 0) if tr else part.sum(0).permute(1, 2, 0)).reshape(n, cin, k, k)
    ref = torch.nn.grad.conv2d_weight(x.double(), (n, cin, k, k), gy.double(), padding=k // 2)
    lib = torch.nn.grad.conv2d_weight(x, (n, cin, k, k), gy, padding=k // 2)
    err = float((got.double() - ref).abs().max() / ref.abs().max()); err_lib = float((lib.double() - ref).abs().max() / ref.abs().max())
    line = "B%d cin%d n%d k%d %dx%d splits %d tr %d status %d  err %.2e (cudnn %.2e)" % (b, cin, n, k, h, w, splits, int(tr), int(status), err, err_lib)
    if time_it:
        def t(f, iters=20):
            for _ in range(3): f()
            torch.cuda.synchronize(); s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(iters): f()
            e.record(); torch.cuda.synchronize(); return s.elapsed_time(e) / iters * 1e3
        mine = t(mine_fn)
        theirs = t(lambda: torch.nn.grad.conv2d_weight(x, (n, cin, k, k), gy, padding=k // 2))
        line += "   %.1f us vs cudnn %.1f us" % (mine, theirs)
    print(line, flush=True)
check(2, 32, 32, 1, 8, 16, False)
check(2, 32, 32, 3, 8, 16, False)
check(4, 24, 40, 3, 16, 16, False)
check(64, 192, 96, 3, 16, 16)
check(64, 192, 192, 1, 16, 16)
check(64, 96, 588, 3, 16, 16)
check(64, 192, 96, 3, 8, 8)
check(64, 192, 192, 1, 8, 8)
check(1, 96, 288, 1, 512, 32)
check(1, 96, 192, 1, 512, 32)
check(64, 320, 160, 3, 16, 16)
check(8, 64, 64, 3, 32, 32)
check(4, 64, 64, 3, 64, 64)
