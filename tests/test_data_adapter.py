"""CPU tests of the real-data adapter (recnet_b200/data.py) against golden vectors produced by the reference's own
dataset/MSVD.py + dataset/transform.py (tests/golden/make_data_golden.py)."""
import json
import os

import numpy as np
import pytest
import torch

import recnet_b200
from recnet_b200 import data as D

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "data_pipeline.json")))


@pytest.fixture
def csv_path(tmp_path):
    p = tmp_path / "captions.csv"
    p.write_text(GOLD["csv"], encoding="utf-8")
    return str(p)


def _vocab(csv_path):
    return D.Vocabulary.from_csv(csv_path, min_count=GOLD["min_count"], caption_max_len=GOLD["caption_max_len"])


def test_vocabulary_matches_reference_msvdvocab(csv_path):
    v = _vocab(csv_path)
    assert v.word2idx == GOLD["word2idx"] and list(v.word2idx) == sorted(GOLD["word2idx"], key=GOLD["word2idx"].get)
    for k in ("n_vocabs", "n_words", "n_vocabs_untrimmed", "n_words_untrimmed", "max_sentence_len"):
        assert getattr(v, k) == GOLD[k], k
    assert v.idx2word[v.word2idx["man"]] == "man"


def test_caption_pipeline_matches_reference_transforms(csv_path):
    v = _vocab(csv_path)
    for p in GOLD["probes"]:
        assert D.sentence_to_words(p["text"], GOLD["caption_max_len"]) == p["words"], p["text"]
        ids = v.encode(p["text"])
        assert ids.dtype == torch.long and ids.tolist() == p["ids"], p["text"]
        assert len(ids) == GOLD["max_sentence_len"] + 1                       # caption_max_len + 1 rows of `targets`
    assert v.decode(v.encode("A man is cooking a fish.")) == "a man is cooking a"     # 'fish' is below min_count: dropped, no <UNK>


@pytest.mark.parametrize("case", GOLD["sampling"], ids=lambda c: f"{c['method']}-{c['n']}")
def test_frame_sampling_matches_reference(case):
    n = case["n"]
    frames = np.stack([np.array([k, k + 0.5], dtype=np.float32) for k in range(n)])
    np.random.seed(case["seed"])                      # the reference draws from numpy's global generator
    got = D.sample_frames(frames, 28, case["method"])
    assert got.shape == (28, 2) and got.dtype == torch.float32
    assert got[:, 0].tolist() == case["first_col"]
    if n < 28:
        assert bool((got[n:] == 0).all())             # ZeroPadIfLessThan


def test_unknown_sampling_method_raises_like_the_reference():
    with pytest.raises(NotImplementedError):
        D.sample_frames(np.zeros((40, 2), dtype=np.float32), 28, "bogus")


def test_dataset_pairs_and_collate(csv_path, tmp_path):
    v = _vocab(csv_path)
    feats = {"vidB_5_9": np.random.rand(40, 6).astype(np.float32), "vidA_0_10": np.random.rand(9, 6).astype(np.float32),
             "vidD_3_8": np.random.rand(28, 6).astype(np.float32)}
    npz = tmp_path / "feats.npz"
    np.savez(npz, **feats)
    for src in (feats, str(npz)):
        ds = D.CaptionFeatureDataset(src, csv_path, v, n_frames=28)
        assert [[vid, c] for vid, _, c in ds.pairs] == GOLD["pairs"]
        vid, f, t = ds[0]
        assert f.shape == (28, 6) and t.shape == (GOLD["max_sentence_len"] + 1,)
    batch = [ds[i] for i in range(3)]
    vids, videos, targets = D.collate(batch, batch_size=5)                     # short batch: padded with copies of the last item
    assert vids[3:] == ["PAD", "PAD"] and videos.shape == (5, 28, 6) and videos.dtype == torch.float32
    assert targets.shape == (GOLD["max_sentence_len"] + 1, 5) and targets.dtype == torch.long and targets.is_contiguous()
    assert torch.equal(targets[:, 4], targets[:, 2]) and torch.equal(videos[3], videos[2])
    masks = targets > D.PAD
    assert int(masks[:, 0].sum()) == int((ds[0][2] > 0).sum())
    # clips shorter than n_frames are zero padded, longer ones sampled uniformly
    assert bool((ds[2][1][9:] == 0).all()) and torch.equal(ds[2][1][:9], torch.from_numpy(feats["vidA_0_10"]))


def test_hdf5_without_h5py_fails_loudly(tmp_path):
    try:
        import h5py  # noqa: F401
        pytest.skip("h5py present")
    except ImportError:
        pass
    with pytest.raises(RuntimeError, match="h5py"):
        D.load_features(str(tmp_path / "feats.hdf5"))
