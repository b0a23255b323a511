"""Debug: sliced vs tile kernels (forward and backward dH) on the failing edge-case graph; run twice with RENET_GATHER_KERNEL."""
import os, sys
import numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import gpu_helpers as G
N, E, R2 = 40000, 17000, 480
rng = np.random.RandomState(N + E)
dst = rng.randint(0, N, E)
dead = rng.rand(N) < 0.7
dead[:7] = True; dead[-9:] = True
alive = np.flatnonzero(~dead)
dst = alive[rng.randint(0, len(alive), E)]
src, et = rng.randint(0, N, E), rng.randint(0, R2, E)
deg = np.bincount(dst, minlength=N).astype(np.float32); deg[deg == 0] = 1
norm = 1.0 / deg
torch.manual_seed(0)
H, W, Wl = torch.randn(N, 200) * 0.3, torch.randn(R2, 400) * 0.1, torch.randn(200, 200) * 0.07
Gout = torch.randn(N, 200)
rp, cs, ct = G.csr_from_coo(src, dst, et, N)
Hd, Wd, Wld, nd = H.to(G.DEV), W.to(G.DEV), Wl.to(G.DEV), G.d(norm)
out = G.layer_fwd(Hd, None, Wd, Wld, rp, cs, ct, nd, N, E, 200, 200, 100, True)
dH, dW, dWl = G.layer_bwd(Hd, None, Wd, Wld, src, dst, et, nd, out, Gout.to(G.DEV), N, E, 200, 200, 100, True)
k = os.environ.get('RENET_GATHER_KERNEL', 'auto')
np.savez('gpurun_out/dbgsl_%s.npz' % k, out=out.cpu().numpy(), dH=dH.cpu().numpy())
other = 'gpurun_out/dbgsl_%s.npz' % ('tile' if k != 'tile' else 'sliced')
if os.path.exists(other):
    o = np.load(other)
    odeg = np.bincount(src, minlength=N)
    rps = np.concatenate(([0], np.cumsum(odeg)))
    for name, a in (('out', out.cpu().numpy()), ('dH', dH.cpu().numpy())):
        d = np.abs(a - o[name]).max(1)
        bad = np.flatnonzero(d > 1e-4)
        print(name, 'rows differing', len(bad), 'of', N, 'max', d.max())
        for r in bad[:12]:
            cols = np.flatnonzero(np.abs(a[r] - o[name][r]) > 1e-4)
            print('   row', r, 'out-deg', odeg[r], 'in-deg', int(deg[r]), 'rp', rps[r], 'cols', cols[:6], '..', cols[-3:], 'n', len(cols), 'got', a[r][cols[:2]], 'ref', o[name][r][cols[:2]])
