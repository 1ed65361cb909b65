import re, sys, torch, numpy as np
sys.path.insert(0, ".")
import phylo2vec_b200 as p2v
dev = torch.device("cuda:0")
for n, B in ((100, 20000), (1000, 4000)):
    k = n - 1
    g = torch.Generator(device=dev).manual_seed(n)
    hi = (2 * torch.arange(k, device=dev) + 1).to(torch.float64)
    V = (torch.rand((B, k), generator=g, device=dev, dtype=torch.float64) * hi).to(torch.int64)
    strs = p2v.to_newick_batch(V)
    bare = [re.sub(r"\)\d+", ")", s) for s in strs]
    enc = [s.encode() for s in bare]
    stride = (max(len(e) for e in enc) + 16) & ~15
    host = np.zeros((B, stride), dtype=np.uint8)
    for b, e in enumerate(enc):
        host[b, :len(e)] = np.frombuffer(e, dtype=np.uint8)
    buf = torch.from_numpy(host).to(dev); ln = torch.tensor([len(e) for e in enc], device=dev)
    # replicate to a larger batch
    rep = 16
    buf = buf.repeat(rep, 1).contiguous(); ln = ln.repeat(rep).contiguous(); Vr = V.repeat(rep, 1)
    f = lambda: p2v.from_newick_batch(buf, lengths=ln, n_leaves=n)
    out = f(); torch.cuda.synchronize()
    a, b2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3): f()
    b2.record(); torch.cuda.synchronize()
    t = a.elapsed_time(b2) / 3e3
    print(f"n={n} no-parent-labels from_newick B={B*rep} {B*rep/t:.3e} trees/s equal={bool(torch.equal(out, Vr))}", flush=True)
