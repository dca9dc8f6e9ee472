"""Helpers shared by the oracle (CPU) and CUDA (GPU) parity tests: load the fixtures that
tests/golden/make_golden.py produced from the reference, rebuild their inputs."""
import os

import numpy as np
import torch

from oracle import nrms_oracle as O
from pytorch_news_recommender_b200 import synthetic as S

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
N_SAMPLES = 48


def sample_idx(n):
    return np.unique(np.linspace(0, n - 1, N_SAMPLES).astype(np.int64))


class Case:
    def __init__(self, name):
        self.name = name
        self.z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        B, T, H, C, D, h, Q, n_news = [int(v) for v in self.z["meta/dims"]]
        self.B, self.T, self.H, self.C, self.D, self.h, self.Q = B, T, H, C, D, h, Q
        self.vocab = int(self.z["meta/vocab"])
        self.seed = int(self.z["meta/seed"])
        self.cfg = O.OracleConfig(T, H, C - 1, D, h, Q, float(self.z["meta/dropout"]), float(self.z["meta/lr"]))
        self.batch = {
            "browsed_titles": torch.from_numpy(self.z["in/browsed_titles"].astype(np.int64)),
            "candidate_titles": torch.from_numpy(self.z["in/candidate_titles"].astype(np.int64)),
            "candidate_mask": torch.from_numpy(self.z["in/candidate_mask"]),
        }
        self.table = S.make_embedding_table(self.vocab, D, seed=self.seed)

    def state_dict(self):
        """Initial weights: stored in full for the tiny case, replayed from seed 42 otherwise
        (and checked against the stored norms/samples)."""
        if f"sd0/{O.TABLE_KEY}" in self.z:
            return {k: torch.from_numpy(self.z[f"sd0/{k}"].copy()) for k in O.state_dict_keys()}
        return O.init_state_dict(self.cfg, self.table, seed=42)

    def masks(self, step):
        """The dropout multipliers make_golden.py injected at train step `step`."""
        rng = np.random.default_rng(1000 + self.seed)
        n_titles = self.B * (self.C + self.H)
        m = None
        for _ in range(step + 1):
            keep = rng.random((2, n_titles, self.T, self.D)) >= self.cfg.dropout
            m = keep.astype(np.float32) / np.float32(1.0 - self.cfg.dropout)
        return torch.from_numpy(m[0]), torch.from_numpy(m[1])

    def summary(self, prefix, key):
        return (float(self.z[f"{prefix}/{key}/norm"]), float(self.z[f"{prefix}/{key}/sum"]),
                self.z[f"{prefix}/{key}/samples"])


def check_summary(case, prefix, tensors, rtol, atol_frac=1e-4, what="", abs_floor=1e-6):
    """Compare tensors against the stored (norm, samples) summaries.  Sample tolerance is
    rtol*|x| + atol_frac*rms(tensor): single entries of a gradient can sit far below its scale.
    abs_floor absorbs tensors that are mathematically zero (the W_K bias gradient: softmax is
    invariant to a per-row constant), where only rounding noise is left."""
    for k, t in tensors.items():
        a = t.detach().cpu().numpy().astype(np.float64).ravel()
        norm, _sum, samples = case.summary(prefix, k)
        got_norm = float(np.sqrt((a * a).sum()))
        assert abs(got_norm - norm) <= rtol * max(norm, 1e-12) + abs_floor * np.sqrt(a.size), \
            f"{what}{prefix}/{k}: norm {got_norm} vs golden {norm}"
        rms = norm / np.sqrt(max(a.size, 1))
        got = a[sample_idx(a.size)]
        tol = rtol * np.abs(samples) + atol_frac * rms + abs_floor
        bad = np.abs(got - samples) > tol
        assert not bad.any(), f"{what}{prefix}/{k}: samples differ: {got[bad][:4]} vs {samples[bad][:4]}"


class BertCase:
    """Fixtures of the `nrms` variant (tests/golden/make_golden_bert.py)."""

    def __init__(self, name):
        from oracle import nrms_bert_oracle as OB
        self.name = name
        self.z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        B, H, C, E, heads, Q, n_news = [int(v) for v in self.z["meta/dims"]]
        self.B, self.H, self.C, self.E, self.heads, self.Q, self.n_news = B, H, C, E, heads, Q, n_news
        self.seed = int(self.z["meta/seed"])
        self.cfg = OB.BertOracleConfig(H, C - 1, E, E, heads, Q, float(self.z["meta/dropout"]), float(self.z["meta/lr"]))
        self.batch = {
            "browsed_ids": torch.from_numpy(self.z["in/browsed_ids"].astype(np.int64)),
            "candidate_ids": torch.from_numpy(self.z["in/candidate_ids"].astype(np.int64)),
            "browsed_mask": torch.from_numpy(self.z["in/browsed_mask"]),
            "candidate_mask": torch.from_numpy(self.z["in/candidate_mask"]),
        }
        self.table = S.make_news_vector_table(n_news, E, seed=self.seed)
        self.keys = OB.state_dict_keys()
        self._OB = OB

    def state_dict(self):
        if f"sd0/{self.keys[0]}" in self.z:
            return {k: torch.from_numpy(self.z[f"sd0/{k}"].copy()) for k in self.keys}
        return self._OB.init_state_dict(self.cfg, self.table, seed=42)

    def mults(self, step):
        """The dropout multipliers make_golden_bert.py injected at train step `step`."""
        rng = np.random.default_rng(2000 + self.seed)
        p = self.cfg.dropout
        out = None
        for _ in range(step + 1):
            out = {}
            for k, shape in (("cand", (self.B, self.C, self.E)), ("hist", (self.B, self.H, self.E)),
                             ("attn", (self.B, self.heads, self.H, self.H))):
                out[k] = torch.from_numpy((rng.random(shape) >= p).astype(np.float32) / np.float32(1.0 - p))
        return out

    def summary(self, prefix, key):
        return (float(self.z[f"{prefix}/{key}/norm"]), float(self.z[f"{prefix}/{key}/sum"]),
                self.z[f"{prefix}/{key}/samples"])
