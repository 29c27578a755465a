"""Eager vs CUDA-graph replay of the eval forward (development helper; small batches are launch-bound in eager mode).
    python tools/graph_bench.py [B ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tfswa_unet_b200 as T

T.set_precision("bf16")
torch.manual_seed(0)
model = T.TFSWAUNet(2, 2, [2, 2, 6, 2], [32, 64, 128, 256], 8, 4, 8).eval().cuda()


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for B in [int(a) for a in sys.argv[1:]] or [1, 2, 8]:
    x = torch.randn(B, 2, 1025, 517, device="cuda")
    with torch.no_grad():
        ref = model(x)
        t_eager = timed(lambda: model(x))
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                model(x)
        torch.cuda.current_stream().wait_stream(s)
        with torch.cuda.graph(g):
            y = model(x)
        t_graph = timed(g.replay)
        g.replay()
        torch.cuda.synchronize()
        same = torch.equal(y, ref)
    print(f"B={B}: eager {t_eager:.2f} ms, graph replay {t_graph:.2f} ms ({t_eager / t_graph:.3f}x), identical output: {same}")
