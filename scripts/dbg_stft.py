import sys, numpy as np, torch
sys.path.insert(0, '.')
from maximilian_b200 import capi, workloads as W
C = int(sys.argv[1]); n, hop = 1024, 512
ctx = capi.Context(0, 48000)
st = capi.Stft(C, n, hop, ctx=ctx); mf = capi.Mfcc(512, 42, 40, 20.0, 20000.0, ctx=ctx)
x = torch.from_numpy(np.tile(W.channel_streams(64, hop, seed=5), (C // 64, 1))).cuda()
co = torch.empty((C, 1, 40), dtype=torch.float64, device='cuda')
for k in range(6):
    f = st.process_device(x.data_ptr(), hop, 1, hop, 1, mfcc=mf, coeffs=co.data_ptr(), stream=0)
    torch.cuda.synchronize(); print(k, f, float(co.abs().sum()), flush=True)
