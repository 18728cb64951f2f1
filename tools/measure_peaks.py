"""Measure the roofline denominators MEASURED_PEAKS.json does not carry (TF32 and FP64 matmul)
with torch.matmul (cuBLAS), next to the bf16 / HBM-copy figures for cross-checking.
Writes gpurun_out/peaks_extra.json."""
import json
import os
import time

import torch


def bench(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) * 1e-3)
    return best


def main():
    out = {'gpu': torch.cuda.get_device_name(0), 'when': time.strftime('%Y-%m-%dT%H:%M:%SZ', time.gmtime())}
    n = 8192
    for name, dtype, tf32 in (('bf16', torch.bfloat16, False), ('tf32', torch.float32, True),
                              ('fp32', torch.float32, False), ('fp64', torch.float64, False)):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        m = n if name != 'fp64' else 4096
        a = torch.randn(m, m, device='cuda', dtype=dtype)
        b = torch.randn(m, m, device='cuda', dtype=dtype)
        t = bench(lambda: torch.matmul(a, b), iters=10 if name != 'fp64' else 5)
        out[name + '_tflops'] = 2 * m ** 3 / t / 1e12
    x = torch.empty(1 << 29, device='cuda', dtype=torch.float32)
    y = torch.empty_like(x)
    t = bench(lambda: y.copy_(x))
    out['hbm_copy_gbs'] = 2 * x.numel() * 4 / t / 1e9
    t = bench(lambda: x.sum())
    out['hbm_read_gbs'] = x.numel() * 4 / t / 1e9
    os.makedirs('gpurun_out', exist_ok=True)
    with open('gpurun_out/peaks_extra.json', 'w') as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
