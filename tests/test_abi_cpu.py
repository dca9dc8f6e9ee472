"""CPU-only checks of the C-ABI boundary: the library builds for sm_100a, loads, exports every
symbol include/nrms_b200.h declares, and its host-side validation returns the documented
error codes (no kernel is launched here)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "nrms_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nrms_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_bound_and_exported(built_lib):
    from pytorch_news_recommender_b200 import _lib
    decl = declared_symbols()
    assert len(decl) >= 20
    assert set(decl) == set(_lib.SIGNATURES.keys()), set(decl) ^ set(_lib.SIGNATURES.keys())
    for name in decl:
        assert hasattr(built_lib, name)
    assert built_lib.nrms_abi_version() == _lib.ABI_VERSION


def test_library_has_no_torch_dependency(built_lib):
    from pytorch_news_recommender_b200 import _lib
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in out and "c10" not in out and "python" not in out.lower()


def test_sass_targets_sm100a(built_lib):
    from pytorch_news_recommender_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_param_count_and_sizes(built_lib):
    from pytorch_news_recommender_b200 import ops
    assert ops.encoder_param_count(300, 200) == 331300          # 2 x 331,300 = 662,600 dense params
    shape = ops.EncoderShape(64 * 55, 30, 300, 10, 200, 70000)
    M = 64 * 55 * 30
    assert ops.saved_bytes(shape) >= 4 * M * (900 + 10 + 300 + 200)
    assert ops.scratch_bytes(shape) >= 4 * M * (300 + 200 + 900)
    assert ops.embedding_plan_bytes(M, 70000) >= 4 * (3 * 70000 + 2 * M)


def test_validation_errors(built_lib):
    from pytorch_news_recommender_b200 import _lib, ops
    from pytorch_news_recommender_b200._lib import EncoderDims, NrmsError
    lib = built_lib
    bad = [
        EncoderDims(0, 30, 300, 10, 200, 100, 0.0, 0, 0),      # n_seq
        EncoderDims(4, 257, 300, 10, 200, 100, 0.0, 0, 0),     # seq_len > 256
        EncoderDims(4, 30, 302, 10, 200, 100, 0.0, 0, 0),      # D % 4
        EncoderDims(4, 30, 300, 7, 200, 100, 0.0, 0, 0),       # D % heads
        EncoderDims(4, 30, 300, 5, 200, 100, 0.0, 0, 0),       # head dim 60 > 32
        EncoderDims(4, 30, 300, 10, 201, 100, 0.0, 0, 0),      # Q % 4
        EncoderDims(4, 30, 300, 10, 200, 100, 1.0, 0, 0),      # dropout_p
        EncoderDims(4, 30, 300, 10, 200, 100, 0.0, 9, 0),      # gemm_mode
    ]
    for d in bad:
        assert lib.nrms_encoder_saved_bytes(d) == -1
        assert len(lib.nrms_last_error()) > 0
    # NULL / misaligned pointers are rejected before any launch
    ok = EncoderDims(4, 30, 300, 10, 200, 100, 0.0, 0, 0)
    rc = lib.nrms_news_encoder_fwd(ok, None, None, None, None, None, 0, None)
    assert rc == -5
    rc = lib.nrms_news_encoder_fwd(ok, 8, 16, 16, 16, 16, 0, None)
    assert rc == -2 and b"aligned" in lib.nrms_last_error()
    assert lib.nrms_adam_step(16, 16, 16, 16, 0, 1, 1e-3, 0.9, 0.999, 1e-8, 1.0, None) == -1
    assert lib.nrms_embedding_plan_bytes(-1, 10) == -1
    with pytest.raises(NrmsError):
        ops.saved_bytes(ops.EncoderShape(4, 300, 300, 10, 200, 10))


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under the package may reference it."""
    pkg = os.path.join(ROOT, "pytorch_news_recommender_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("nrms_oracle_free", ""), os.path.join(dirpath, f)


def test_model_requires_cuda(built_lib, tmp_path):
    """No CPU fallback: on a CPU-only host the model constructs but forward raises."""
    import torch

    from pytorch_news_recommender_b200 import synthetic as S
    from pytorch_news_recommender_b200._lib import NrmsError
    from pytorch_news_recommender_b200.config import Config
    from pytorch_news_recommender_b200.model import NRMS_V0
    cfg = Config("T")
    with pytest.raises(AttributeError):
        NRMS_V0(cfg)                                            # __nrms__() not called yet
    cfg.__nrms__()
    cfg.n_words_title, cfg.history_len, cfg.sample_size = 6, 5, 2
    S.save_embedding_npz(str(tmp_path / "e.npz"), S.make_embedding_table(50, 300, 0))
    cfg.data_path, cfg.word_embedding_pretrained, cfg.device = str(tmp_path) + "/", "e.npz", torch.device("cpu")
    model = NRMS_V0(cfg)
    assert len(model.state_dict()) == 19
    pool = S.make_news_pool(20, 6, 50)
    batch = S.make_train_batch(pool, 2, 5, 2)
    assert batch["browsed_titles"].dtype == torch.int64 and batch["candidate_mask"].dtype == torch.uint8
    with pytest.raises(NrmsError):
        model(batch)
