import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from oracle import xt_oracle as orc
from xingtian_b200 import capi
from xingtian_b200.engine import Net
lib = capi.lib()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
arch = orc.ppo_cnn_arch()
w = orc.init_weights(arch, seed=11)
for k in w:
    if k.endswith("/bias"):
        w[k] = (np.random.default_rng(1).standard_normal(w[k].shape) * 0.05).astype(np.float32)
rng = np.random.default_rng(2)
nrows = B + 3
obs = rng.integers(0, 256, (nrows,) + arch["state_dim"], dtype=np.uint8)
USE_IDX = os.environ.get("USE_IDX", "1") == "1"
idx = rng.permutation(nrows)[:B].astype(np.int32) if USE_IDX else np.arange(B, dtype=np.int32)
idx_d = torch.from_numpy(idx).cuda()
obs_d = torch.from_numpy(obs).cuda()
gh = {h: rng.standard_normal((B, n)).astype(np.float32) for h, n in (("pi_latent", 4), ("output_value", 1))}
res = {}
for mode in (0, 1):
    lib.xtb_set_tc_mode(mode)
    net = Net(arch, max_batch=B)
    net.set_weights(w)
    net.forward(obs_d, B, idx=idx_d if USE_IDX else None)
    for h in gh:
        net.tensor_grad(h)[:B].copy_(torch.from_numpy(gh[h]).cuda())
    net.backward(obs_d, B, list(gh.keys()), idx=idx_d if USE_IDX else None)
    torch.cuda.synchronize()
    res[mode] = {n: (net.tensor(n)[:B].cpu().numpy().copy(), net.tensor_grad(n)[:B].cpu().numpy().copy()) for n, _, _, _ in arch["layers"]}
    res[mode]["_g"] = net.get_weights(net.grads)
shapes = orc.tensor_shapes(arch)
for n, _, _, _ in arch["layers"]:
    o0, g0 = res[0][n]; o1, g1 = res[1][n]
    do = np.abs(o0 - o1).max() / max(np.abs(o0).max(), 1e-9)
    dg = np.abs(g0 - g1); dgm = dg.max() / max(np.abs(g0).max(), 1e-9)
    print(n, "out diff %.2e  gout diff %.2e" % (do, dgm))
    if dgm > 1e-4:
        shp = (B,) + shapes[n]
        d = dg.reshape(shp)
        bad = np.argwhere(d > 1e-4 * np.abs(g0).max())
        print("   bad count", len(bad), "of", d.size)
        print("   bad samples:", np.unique(bad[:, 0])[:20], " ys:", np.unique(bad[:, 1]), " xs:", np.unique(bad[:, 2]) if len(shp) > 2 else "")
        if len(shp) == 4:
            print("   bad channels:", np.unique(bad[:, 3]))
        print("   first few:", bad[:6].tolist(), [float(g0.reshape(shp)[tuple(i)]) for i in bad[:3]], [float(g1.reshape(shp)[tuple(i)]) for i in bad[:3]])
for k in res[0]["_g"]:
    a, b = res[0]["_g"][k], res[1]["_g"][k]
    print(k, "%.2e" % (np.abs(a - b).max() / max(np.abs(a).max(), 1e-12)))

params = {k: torch.from_numpy(v.copy()).requires_grad_(True) for k, v in w.items()}
ref = orc.forward(arch, params, obs[idx], keep=True)
loss = sum((ref[h] * torch.from_numpy(gh[h])).sum() for h in gh)
loss.backward()
for mode in (0, 1):
    print("mode", mode, {k: "%.1e" % (np.abs(res[mode]["_g"][k] - params[k].grad.numpy()).max() / np.abs(params[k].grad.numpy()).max()) for k in w})
