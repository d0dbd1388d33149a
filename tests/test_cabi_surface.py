"""CPU-side checks of the drop-in boundary: the shared library loads without a GPU or a CUDA driver,
exports every symbol include/tfkaldi_b200.h declares, and refuses to compute without a device."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from tfkaldi_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__

        __graft_entry__.build()
    return _lib.load()


def header_symbols():
    text = open(os.path.join(ROOT, "include", "tfkaldi_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tfk_[a-z_0-9]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from tfkaldi_b200 import _lib

    declared = header_symbols()
    assert len(declared) >= 25
    bound = {name for name, _, _ in _lib.SIGNATURES}
    for name in declared:
        assert hasattr(lib, name), "library does not export %s" % name
        assert name in bound, "ctypes binding lacks %s" % name
    assert bound <= set(declared), bound - set(declared)


def test_abi_version_and_config_defaults(lib):
    from tfkaldi_b200 import _lib

    assert lib.tfk_abi_version() == _lib.TFK_ABI_VERSION
    cfg = _lib.TfkConfig()
    lib.tfk_default_config(C.byref(cfg))
    assert (cfg.num_layers, cfg.input_dim, cfg.hidden_dim, cfg.output_dim) == (6, 440, 2048, 1936)
    assert abs(cfg.bn_eps - 1e-3) < 1e-9 and abs(cfg.bn_decay - 0.999) < 1e-7  # tf.contrib.layers.batch_norm
    assert abs(cfg.adam_beta1 - 0.9) < 1e-7 and abs(cfg.adam_beta2 - 0.999) < 1e-7 and abs(cfg.adam_eps - 1e-8) < 1e-12
    assert cfg.keep_prob == 1.0 and cfg.precision == _lib.TFK_PREC_BF16


def test_no_silent_cpu_fallback(lib):
    """without a CUDA device tfk_create must fail loudly (TFK_ECUDA), never compute on the host"""
    import torch

    from tfkaldi_b200 import _lib

    if torch.cuda.is_available():
        pytest.skip("a GPU is visible: the loud-failure path is exercised on the CPU box")
    assert lib.tfk_device_count() == 0
    cfg = _lib.TfkConfig()
    lib.tfk_default_config(C.byref(cfg))
    h = C.c_void_p()
    assert lib.tfk_create(C.byref(cfg), C.byref(h)) == _lib.TFK_ECUDA
    assert b"no CUDA device" in lib.tfk_last_error(None)
    from tfkaldi_b200.engine import Engine

    with pytest.raises(RuntimeError):
        Engine(2, 40, 32, 10, 64)


def test_bad_configs_are_rejected(lib):
    from tfkaldi_b200 import _lib

    cfg = _lib.TfkConfig()
    lib.tfk_default_config(C.byref(cfg))
    h = C.c_void_p()
    cfg.abi_version = 99
    assert lib.tfk_create(C.byref(cfg), C.byref(h)) == _lib.TFK_EINVAL
    lib.tfk_default_config(C.byref(cfg))
    cfg.keep_prob = 0.0  # Dropout asserts 0 < keep <= 1 (classifiers/activation.py:127)
    assert lib.tfk_create(C.byref(cfg), C.byref(h)) == _lib.TFK_EINVAL
    lib.tfk_default_config(C.byref(cfg))
    cfg.nonlin = 7  # 'unkown nonlinearity' (nnet.py:65)
    assert lib.tfk_create(C.byref(cfg), C.byref(h)) == _lib.TFK_EINVAL


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under tfkaldi_b200/ may import, include or execute it
    (comments may mention it)"""
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|#\s*include\s*[<\"].*oracle|oracle[/.]\w+\.(py|so|c)\b|import_module\(.*oracle", re.M)
    for base, _, files in os.walk(os.path.join(ROOT, "tfkaldi_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, f)).read()
                assert not pat.search(text), os.path.join(base, f)


def test_work_list_scheduler_covers_every_tile_exactly_once():
    """host half of the CTA-pair GEMM (gemm_schedule_tile_lists): `selftest_gemm sched` sweeps 2352 launch shapes
    (forward, fused wgrad+dgrad, split-K wgrad; ragged N; bf16 / bf16x3) and checks coverage, termination and determinism
    of the per-pair work lists — no GPU involved."""
    import subprocess

    binary = os.path.join(ROOT, "tfkaldi_b200", "csrc", "build", "selftest_gemm")
    if not os.path.exists(binary):
        import __graft_entry__

        __graft_entry__.build()
    out = subprocess.run([binary, "sched"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and "ALL PASS" in out.stdout and "2352 launch shapes" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
