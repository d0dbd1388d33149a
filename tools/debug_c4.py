"""debug aid: per-tensor gradient errors of one full-size C4 step (bf16x3) against the oracle"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from test_gpu_zfullsize import make, C4
from oracle.dnn_oracle import set_matmul_backend
from tfkaldi_b200 import _lib as L

set_matmul_backend("torch")
nonlin = sys.argv[1] if len(sys.argv) > 1 else "linear"
B = 4096
orc, eng, rng, cfg, _ = make(dict(C4, nonlin=nonlin), B, "bf16x3", seed=43)
x = rng.standard_normal((B, 440)).astype(np.float32)
y = rng.integers(0, 3401, B)
eng.set_dropout_seed(4321)
eng.accumulate(x, y)
orc.accumulate(x, y, dropout_seed=4321)
print("loss", eng.get_scalar(L.S_LOSS_SUM), orc.loss_sum)
kinds = {"W": L.T_GRAD_W, "b": L.T_GRAD_B, "beta": L.T_GRAD_BETA}
for k, want in orc.grads.items():
    stem = k.rstrip("0123456789"); layer = int(k[len(stem):])
    if stem == "b" and layer < 6: continue
    got = eng.get_tensor(kinds[stem], layer).astype(np.float64); want = want.astype(np.float64)
    print(nonlin, os.environ.get("TFK_SOFTMAX"), os.environ.get("TFK_BN_FROM_Y"), k, "max %.2e l2 %.2e" % (np.abs(got-want).max()/np.abs(want).max(), np.linalg.norm(got-want)/np.linalg.norm(want)))
# the masks
for l in range(6):
    act = eng.activation(l, B).cpu().numpy()
    _, caches = orc.forward(x, training=True, dropout_seed=4321) if l == 0 else (None, caches)
    kept_g = act != 0
    kept_o = caches[l].y != 0
    print("layer", l, "mask mismatches", int((kept_g != kept_o).sum()), "max |y diff|", float(np.abs(act - caches[l].y).max()))
