"""Experiment helper: times the streaming STFT kernel on configs[3] with and without its MFCC stage (device pointers, CUDA events)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from maximilian_b200 import capi

dev = torch.device("cuda:0"); torch.cuda.set_device(dev)
ctx = capi.Context(0, 48000)
C, n, hop, H = 65536, 1024, 512, 8
rng = np.random.default_rng(5)
x = torch.from_numpy(np.tile(rng.standard_normal((1024, H * hop)).astype(np.float32) * 0.3, (C // 1024, 1))).to(dev)
stream = torch.cuda.Stream()
out = {}
for name in ("mfcc", "fft_only"):
    st = capi.Stft(C, n, hop, ctx=ctx)
    mf = capi.Mfcc(n // 2, 42, 40, 20.0, 20000.0, ctx=ctx) if name == "mfcc" else None
    co = torch.empty((C, H, 40), dtype=torch.float64, device=dev) if mf else None
    def step():
        st.process_device(x.data_ptr(), H * hop, 1, H * hop, H, mfcc=mf, coeffs=co.data_ptr() if mf else None, stream=stream.cuda_stream)
    for _ in range(5): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(30): step()
    e1.record(stream); torch.cuda.synchronize()
    out[name] = e0.elapsed_time(e1) / 30
print("stft_probe", os.environ.get("MXB_LIB_PATH", "base").split("_")[-1], json.dumps(out))
