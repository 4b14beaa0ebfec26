"""Does tcgen05 kind::tf32 ignore (truncate) the 13 low mantissa bits of an fp32 operand?  If so `hi` can be the raw
fp32 tensor and only `lo = rna(x - trunc(x))` must be materialised."""
import sys, torch
sys.path.insert(0, '/root/repo')
import flowk
from flowk import tc
dev = torch.device('cuda:0')
torch.manual_seed(0)
M, K, N = 1024, 96, 96
a = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) * 0.1
ref = a.double() @ w.double().t()
def run(a_hi, a_lo, w_hi, w_lo):
    y = torch.empty(M, N, device=dev)
    tc.conv_gemm(a_hi, a_lo, w_hi, w_lo, M // 128, 8, 16, K, N, 1, tc.PRE_BIAS, tc.OUT_F32, out_f32=y)
    torch.cuda.synchronize()
    return float((y.double() - ref).abs().max() / ref.abs().max())
def rna(v): return ((v.view(torch.int32) + 0x1000) & -8192).view(torch.float32)
def trunc(v): return (v.view(torch.int32) & -8192).view(torch.float32)
ah, al = tc.split_hilo(a); wh, wl = tc.split_hilo(w)
print("rna split            ", run(ah, al, wh, wl))
print("raw hi, lo=rna(x-trunc)", run(a.clone(), rna(a - trunc(a)), w.clone(), rna(w - trunc(w))))
print("raw hi, lo=rna(x-rna(x)) (wrong if hw truncates)", run(a.clone(), al, w.clone(), wl))
z = torch.zeros_like(a); zw = torch.zeros_like(w)
print("single pass rna hi   ", run(ah, z, wh, zw))
print("single pass raw      ", run(a.clone(), z, w.clone(), zw))
print("single pass trunc    ", run(trunc(a), z, trunc(w), zw))
