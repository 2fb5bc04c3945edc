"""Round-trip parity + first timing on a GPU box (run under gpurun)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import aacjs_b200 as A
from oracle import oracle as O

rng = np.random.default_rng(7)

def legal_seq(T, rng):
    seq, cur = [], 0
    for _ in range(T):
        if cur in (0, 3): nxt = rng.choice([0, 0, 1])
        elif cur == 1: nxt = 2 if rng.random() < 0.7 else 3
        else: nxt = rng.choice([2, 3])
        seq.append(int(nxt)); cur = nxt
    return seq

def mk_tns(info, rng):
    S, T, Cn = info.shape
    coefs = [0.0, -0.2079117, -0.40673664, -0.58778524, 0.67369562, 0.52643216, 0.36124167, 0.18374951]
    blocks = []
    for fi in info.reshape(-1):
        if not fi['tns_present']:
            blocks.append(None); continue
        short = fi['window_sequence'] == 2
        nw = 8 if short else 1
        nf = [int(rng.integers(0, (2 if short else 4))) if w < nw else 0 for w in range(8)]
        b = bytearray(nf)
        for w in range(8):
            for f in range(nf[w]):
                order = int(rng.integers(0, (8 if short else 21)))
                b += bytes([int(rng.integers(1, (15 if short else 40))), order, int(rng.integers(0, 2)), 0])
                b += np.array([coefs[int(rng.integers(0, 8))] for _ in range(order)], np.float32).tobytes()
        blocks.append(bytes(b))
    return A.pack_tns(blocks)

worst = 0
for case, (S, T, Cn, tns) in enumerate([(1, 1, 1, 0), (2, 5, 2, 0), (3, 7, 2, 0), (2, 9, 1, 0), (1, 6, 3, 0), (2, 6, 2, 1),
                                         (2, 6, 2, 2), (3, 4, 5, 1), (64, 40, 2, 0), (37, 33, 2, 1)]):
    spec = (rng.standard_normal((S, T, Cn, 1024)) * 1e5).astype(np.float32)
    info = np.zeros((S, T, Cn), O.INFO_DTYPE)
    for s in range(S):
        for c in range(Cn):
            info['window_sequence'][s, :, c] = legal_seq(T, rng) if case > 0 else [0]
    info['shape_cur'] = rng.integers(0, 2, (S, T, Cn)); info['shape_prev'] = rng.integers(0, 2, (S, T, Cn))
    info['max_sfb'] = rng.integers(0, 52, (S, T, Cn))
    blob = offs = None
    if tns:
        info['tns_present'] = rng.integers(0, 2, (S, T, Cn)); blob, offs = mk_tns(info, rng)
    ov0 = (rng.standard_normal((S, Cn, 1024)) * 0.5 * 32768).astype(np.float32)
    ovo = ov0.copy(); ref, _ = O.process(spec, info, blob, offs, ovo, sample_index=4, flags=tns, n_threads=8)
    ctx = A.Context(S, Cn, 4, tns)
    ctx.set_overlap(ov0)
    got = ctx.process(spec, info, blob, offs)
    ovg = ctx.get_overlap()
    e = float(np.abs(got.astype(np.float64) - ref).max()); eo = float(np.abs(ovg - ovo).max() / 32768)
    # second call continues the streams
    ref2, _ = O.process(spec, info, blob, offs, ovo, sample_index=4, flags=tns, n_threads=8)
    got2 = ctx.process(spec, info, blob, offs)
    e2 = float(np.abs(got2.astype(np.float64) - ref2).max())
    print(case, (S, T, Cn, tns), 'pcm', e, 'ovl', eo, 'second call', e2, 'launches', ctx.launches, flush=True)
    worst = max(worst, e, e2, eo)
    ctx.close()
print('WORST', worst)

# timing: config 2, device resident
import torch
S, T, Cn = 256, 256, 2
dev = torch.device('cuda:0')
spec = (torch.randn((S, T, Cn, 1024), device=dev) * 3e5)
info = torch.zeros((S, T, Cn, 8), dtype=torch.uint8, device=dev)
info[..., 2] = (torch.arange(S, device=dev) & 1).to(torch.uint8)[:, None, None]
pcm = torch.empty((S, T, 1024, Cn), device=dev)
ctx = A.Context(S, Cn, 4, 0)
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    ctx.process_device(spec.data_ptr(), info.data_ptr(), pcm.data_ptr(), T, st)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
ev[0].record()
for i in range(10):
    ctx.process_device(spec.data_ptr(), info.data_ptr(), pcm.data_ptr(), T, st)
    ev[i + 1].record()
torch.cuda.synchronize()
ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(10)]
ms = float(np.median(ts))
byts = S * T * Cn * 8192
print('config2 ms', ts, 'median', ms, 'GB/s', byts / ms / 1e6, 'frames/s', S * T / ms * 1e3)
print('pcm stats', float(pcm.abs().max()), float(pcm.std()))
