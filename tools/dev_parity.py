"""Developer diagnostic (GPU): step-by-step error growth of the engine against the oracle."""
import math, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle.dnn_oracle import OracleConfig, OracleDNN, reference_init
from tfkaldi_b200.engine import Engine
from tfkaldi_b200 import _lib as L

def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)), float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))

def run(precision, bn=False, keep=1.0, steps=6, lr=1e-3):
    cfg = OracleConfig(num_layers=2, input_dim=440, hidden_dim=256, output_dim=183, batch_norm=bn, keep_prob=keep)
    rng = np.random.default_rng(3)
    params = reference_init(cfg, rng)
    params["W2"] = (rng.standard_normal((256, 183)) / 16).astype(np.float32)
    for l in range(3):
        params[f"b{l}"] = (0.1 * rng.standard_normal(params[f"b{l}"].shape)).astype(np.float32)
    eng = Engine(2, 440, 256, 183, 256, batch_norm=bn, keep_prob=keep, precision=precision)
    eng.load_params(params)
    orc = OracleDNN(cfg, params)
    xe = rng.standard_normal((200, 440)).astype(np.float32)
    prior = np.full(183, 1.0 / 183, np.float32)
    print(f"--- precision={precision} bn={bn} keep={keep}")
    print("step -1 loglik max-rel/l2-rel", rel(eng.loglik(xe, prior).cpu().numpy(), orc.loglik(xe, prior)))
    seed = 100
    for step in range(steps):
        x = rng.standard_normal((256, 440)).astype(np.float32)
        y = rng.integers(0, 183, 256)
        eng.set_dropout_seed(seed); eng.accumulate(x, y); orc.accumulate(x, y, dropout_seed=seed); seed += 3
        g = {f"W{l}": eng.get_tensor(L.T_GRAD_W, l) for l in range(3)}
        gb = {f"b{l}": eng.get_tensor(L.T_GRAD_B, l) for l in range(3)}
        msg = " ".join(f"dW{l}:{rel(g[f'W{l}'], orc.grads[f'W{l}'])[0]:.1e}" for l in range(3))
        msg += " " + " ".join(f"db{l}:{rel(gb[f'b{l}'], orc.grads[f'b{l}'])[0]:.1e}" for l in range(3))
        lg, lo = eng.apply(lr), orc.apply(lr)
        p = eng.dump_params()
        msg += " | " + " ".join(f"W{l}:{np.abs(p[f'W{l}']-orc.p[f'W{l}']).max():.1e}" for l in range(3))
        msg += " " + " ".join(f"b{l}:{np.abs(p[f'b{l}']-orc.p[f'b{l}']).max():.1e}" for l in range(3))
        ll = rel(eng.loglik(xe, prior).cpu().numpy(), orc.loglik(xe, prior))
        print(f"step {step} loss {lg:.6f} vs {lo:.6f} | {msg} | loglik {ll[0]:.1e}/{ll[1]:.1e}")
        ma = eng.get_tensor(L.T_ADAM_M_W, 2); va = eng.get_tensor(L.T_ADAM_V_W, 2)
        print("      adam m/v W2 rel:", rel(ma, orc.m["W2"]), rel(va, orc.v["W2"]))

if __name__ == "__main__":
    run("bf16x3")
    run("bf16x3", bn=True)
    run("bf16")
