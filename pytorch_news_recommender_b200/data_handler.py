"""Batch assembly for the NRMS path: the reference's `data_handler.py` interface
(`load_dataset`, `get_Words_Infos`, `get_Demo_Words_Infos`, `MyDataset`; run_demo.py:32-55) plus
`DeviceBatcher`, the same assembly done on the GPU.

On-disk inputs (the files the reference reads; synthetic writers in `synthetic.write_demo_files`):
  * `<data_path>idx_<file>`            pickled list of samples
        [history_idx, categ_idx, subcateg_idx, imp_idx, imp_categ_idx, imp_subcateg_idx]
        (data_handler.py:100-103; ids are news-table row + 1, 0 is never stored);
  * `<data_path>[demo_]news_title.pkl` dict  news row -> list of n_words_title token ids;
  * `<data_path>[demo_]news_abst.pkl`  dict  news row -> list of n_words_abst token ids.
Building `idx_*.pkl` from the raw MIND dumps (`News_Processor`, nltk tokenisation) is the
reference's preprocessing, out of scope here: `load_dataset` reads the cached list only.

`MyDataset` produces per sample what data_handler.py:185-250 produces (zero int64 arrays, front-aligned
fill, `[:sample_size]` truncation of the candidates, uint8 masks) from dense token tables built once, and
is the host-side mirror that `torch.utils.data.DataLoader` can drive exactly like the reference's.

`DeviceBatcher` replaces DataLoader + MyDataset for the hot path: the ragged sample lists are packed
ONCE into padded id matrices, the title dict becomes a resident `[n_news, T]` int64 table in HBM,
and a batch is ONE launch of our `nrms_assemble_batch` kernel (a warp per sample slot copies the
news id, the mask byte and the title row of that news, id 0 -> the all-zero title).  The ~110 Python dict
look-ups per sample of the reference loader disappear and the batch never touches the host.
The tensors it yields are bit-identical to `default_collate` over `MyDataset` (tests/test_gpu_data.py).
"""
from __future__ import annotations

import os
import pickle
from ast import literal_eval
from typing import Dict, Iterator, List, Optional, Sequence

import numpy as np
import torch
from torch.utils.data import Dataset

from . import ops
from ._lib import NrmsError


def load_dataset(config, file, path, _type=0):
    """data_handler.py:43-49: the cached sample list `path + 'idx_' + file`."""
    cached = path + 'idx_' + file
    if os.path.exists(cached):
        with open(cached, 'rb') as f:
            return pickle.load(f)
    raise FileNotFoundError(
        f"{cached} not found: building it from the raw MIND dumps is the reference's preprocessing "
        "(data_handler.py:52-108, News_Processor), which this package does not rebuild")


def _words_infos(config, prefix: str):
    tp, ap = config.data_path + prefix + 'news_title.pkl', config.data_path + prefix + 'news_abst.pkl'
    if os.path.exists(tp):
        with open(tp, 'rb') as f:
            title_dict = pickle.load(f)
        abst_dict = {}
        if os.path.exists(ap):
            with open(ap, 'rb') as f:
                abst_dict = pickle.load(f)
        return title_dict, abst_dict
    # data_handler.py:119-135: news_words.csv, no header, columns news_id,title,abstract with the
    # token lists as Python literals; the dicts are keyed by the ROW number
    import pandas as pd
    news_df = pd.read_csv(config.data_path + prefix + 'news_words.csv', header=None)
    news_df.columns = ['news_id', 'title', 'abstract']
    title_dict = {i: literal_eval(row['title']) for i, row in news_df.iterrows()}
    abst_dict = {i: literal_eval(row['abstract']) for i, row in news_df.iterrows()}
    with open(tp, 'wb') as f:
        pickle.dump(title_dict, f)
    with open(ap, 'wb') as f:
        pickle.dump(abst_dict, f)
    return title_dict, abst_dict


def get_Words_Infos(config):
    """data_handler.py:113-135."""
    return _words_infos(config, '')


def get_Demo_Words_Infos(config):
    """data_handler.py:137-159."""
    return _words_infos(config, 'demo_')


def _front_aligned(values, slots: int) -> np.ndarray:
    """`values` at the front of a zero int64 vector of `slots` entries (a longer list is a ValueError, the
    broadcast error the reference's slice assignment raises)."""
    out = np.zeros(slots, dtype=np.int64)
    out[:len(values)] = np.asarray(values, dtype=np.int64)
    return out


class _TokenTable:
    """A news-row -> token-list dict as one dense matrix (row r = dict[r]) plus a presence mask, so that a
    sample's titles are ONE fancy-indexing gather instead of a Python list per news."""

    def __init__(self, mapping: Dict[int, Sequence[int]], width: int):
        n = (max(mapping) + 1) if mapping else 0
        self.rows = np.zeros((n, width), dtype=np.int64)
        self.present = np.zeros(n, dtype=bool)
        for r, toks in mapping.items():
            self.rows[r, :] = np.asarray(toks, dtype=np.int64)
            self.present[r] = True

    def lookup(self, news_ids, slots: int) -> np.ndarray:
        """[slots, width]: row k = tokens of news id news_ids[k] (id = row + 1), zero rows behind."""
        out = np.zeros((slots, self.rows.shape[1]), dtype=np.int64)
        r = np.asarray(news_ids, dtype=np.int64) - 1
        if r.size:
            bad = (r < 0) | (r >= self.present.size)
            if bad.any() or not self.present[r].all():
                raise KeyError(int(r[bad][0]) if bad.any() else int(r[~self.present[r]][0]))   # the reference: dict KeyError
            out[:r.size] = self.rows[r]
        return out


class MyDataset(Dataset):
    """Host-side loader with the reference's interface and per-sample output (data_handler.py:161-250:
    same keys, dtypes and shapes, front-aligned zero padding, `[:sample_size]` truncation of the
    candidates, uint8 masks), so that `torch.utils.data.DataLoader` drives it exactly like the
    reference's.  It exists for the drop-in scripts and as the checker of `DeviceBatcher`; the hot path
    assembles batches on the GPU.  `datas` is the list `load_dataset` returns."""

    def __init__(self, config, datas, type=0, words_infos=None):
        super().__init__()
        self.config = config
        self.data_type = type
        self.bacthes = datas          # (sic) the reference's attribute name
        if words_infos is not None:
            self.id2title_dict, self.id2abst_dict = words_infos
        elif getattr(config, 'mode', None) == 'demo':
            self.id2title_dict, self.id2abst_dict = get_Demo_Words_Infos(config)
        else:
            self.id2title_dict, self.id2abst_dict = get_Words_Infos(config)
        self.sample_size = self.config.sample_size + 1 if type < 1 else self.config.max_candidate_size
        self._titles = _TokenTable(self.id2title_dict, config.n_words_title)
        n_abst = getattr(config, 'n_words_abst', 0)
        self._absts = _TokenTable(self.id2abst_dict, n_abst) if (n_abst and self.id2abst_dict) else None

    def __len__(self):
        return len(self.bacthes)

    def __getitem__(self, index):
        H, S = self.config.history_len, self.sample_size
        hist, hist_categ, hist_subcateg, cand, cand_categ, cand_subcateg = self.bacthes[index][:6]
        cand = cand[:S]
        out = {'browsed_lens': len(hist)}
        for side, ids, slots, categ, subcateg in (('browsed', hist, H, hist_categ, hist_subcateg),
                                                  ('candidate', cand, S, cand_categ, cand_subcateg)):
            out[side + '_ids'] = _front_aligned(ids, slots)
            out[side + '_titles'] = self._titles.lookup(ids, slots)
            out[side + '_categ_ids'] = _front_aligned(categ, slots)
            out[side + '_subcateg_ids'] = _front_aligned(subcateg, slots)
            out[side + '_mask'] = torch.from_numpy((np.arange(slots) < len(ids)).astype(np.uint8))
            if self._absts is not None:
                out[side + '_absts'] = self._absts.lookup(ids, slots)
        return out


def pack_samples(datas: Sequence[Sequence[Sequence[int]]], history_len: int, sample_size: int):
    """Ragged sample lists -> (browsed_ids [N,H], browsed_lens [N], candidate_ids [N,S],
    candidate_lens [N]) int64, with MyDataset's semantics: front-aligned, zero padded.  A history
    longer than `history_len` or a candidate list longer than `sample_size` is a ValueError in the
    reference (data_handler.py:206,230 broadcast the lists into fixed-size slices) and here."""
    N = len(datas)
    b_len = np.fromiter((len(d[0]) for d in datas), dtype=np.int64, count=N)
    if N and b_len.max() > history_len:
        raise ValueError(f"a sample has {int(b_len.max())} history items > history_len={history_len}")
    c_len = np.fromiter((len(d[3]) for d in datas), dtype=np.int64, count=N)
    if N and c_len.max() > sample_size:
        raise ValueError(f"a sample has {int(c_len.max())} candidates > {sample_size} slots")
    browsed = np.zeros((N, history_len), dtype=np.int64)
    cand = np.zeros((N, sample_size), dtype=np.int64)
    # one flat copy per matrix instead of a Python-level fill per sample
    if N:
        flat_b = np.fromiter((v for d in datas for v in d[0]), dtype=np.int64, count=int(b_len.sum()))
        rows = np.repeat(np.arange(N), b_len)
        cols = np.arange(int(b_len.sum())) - np.repeat(np.cumsum(b_len) - b_len, b_len)
        browsed[rows, cols] = flat_b
        flat_c = np.fromiter((v for d in datas for v in d[3][:sample_size]), dtype=np.int64, count=int(c_len.sum()))
        rows = np.repeat(np.arange(N), c_len)
        cols = np.arange(int(c_len.sum())) - np.repeat(np.cumsum(c_len) - c_len, c_len)
        cand[rows, cols] = flat_c
    return browsed, b_len, cand, c_len


def title_table_from_dict(id2title_dict: Dict[int, Sequence[int]], n_words_title: int) -> np.ndarray:
    """[n_news, T] int64, row r = id2title_dict[r] (news id r + 1)."""
    n = (max(id2title_dict) + 1) if id2title_dict else 0
    table = np.zeros((n, n_words_title), dtype=np.int64)
    for r, toks in id2title_dict.items():
        table[r, :] = np.asarray(toks, dtype=np.int64)
    return table


class DeviceBatcher:
    """GPU batch assembly with `MyDataset` + `DataLoader(default_collate)` semantics (f1).

    Iterating yields dicts of CUDA tensors with the keys the NRMS path consumes:
    browsed_lens [B] i64, browsed_ids [B,H] i64, browsed_titles [B,H,T] i64, browsed_mask [B,H] u8,
    candidate_ids [B,S] i64, candidate_titles [B,S,T] i64, candidate_mask [B,S] u8.
    (Abstract / category arrays are not read by NRMS_V0 and are not produced.)
    shuffle=True draws one permutation per epoch from a torch.Generator seeded with `seed`.
    """

    def __init__(self, config, datas, type=0, batch_size: Optional[int] = None, shuffle: bool = False,
                 drop_last: bool = False, seed: int = 0, device=None, words_infos=None):
        dev = torch.device(device if device is not None else config.device)
        if dev.type != 'cuda':
            raise NrmsError("DeviceBatcher assembles batches on a CUDA device (use MyDataset + DataLoader on the host)")
        self.device = dev
        self.config = config
        self.sample_size = config.sample_size + 1 if type < 1 else config.max_candidate_size
        self.batch_size = int(batch_size if batch_size is not None else config.batch_size)
        self.shuffle, self.drop_last = shuffle, drop_last
        if words_infos is not None:
            title_dict = words_infos[0]
        elif getattr(config, 'mode', None) == 'demo':
            title_dict = get_Demo_Words_Infos(config)[0]
        else:
            title_dict = get_Words_Infos(config)[0]
        b, bl, c, cl = pack_samples(datas, config.history_len, self.sample_size)
        n_news = (max(title_dict) + 1) if title_dict else 0
        if (b.size and b.max() > n_news) or (c.size and c.max() > n_news):
            raise KeyError("a sample references a news id beyond the title table")   # MyDataset: KeyError too
        self.n = b.shape[0]
        self.titles = torch.from_numpy(title_table_from_dict(title_dict, config.n_words_title)).to(dev)
        self.browsed_ids = torch.from_numpy(b).to(dev)
        self.browsed_lens = torch.from_numpy(bl).to(dev)
        self.candidate_ids = torch.from_numpy(c).to(dev)
        self.candidate_lens = torch.from_numpy(cl).to(dev)
        self._gen = torch.Generator()
        self._gen.manual_seed(seed)

    def __len__(self):
        return self.n // self.batch_size if self.drop_last else (self.n + self.batch_size - 1) // self.batch_size

    def batch(self, index: torch.Tensor) -> Dict[str, torch.Tensor]:
        """The batch of the samples `index` (int64, CUDA)."""
        return ops.assemble_batch(index, self.browsed_ids, self.browsed_lens, self.candidate_ids,
                                  self.candidate_lens, self.titles)

    def __iter__(self) -> Iterator[Dict[str, torch.Tensor]]:
        order = torch.randperm(self.n, generator=self._gen) if self.shuffle else torch.arange(self.n)
        order = order.to(self.device)
        stop = self.n - self.n % self.batch_size if self.drop_last else self.n
        for i in range(0, stop, self.batch_size):
            yield self.batch(order[i:min(i + self.batch_size, stop)].contiguous())
