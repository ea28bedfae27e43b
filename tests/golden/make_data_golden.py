"""Golden vectors for the real-data adapter (recnet_b200/data.py), produced by running the REFERENCE's own vocabulary / caption /
frame-sampling code (dataset/MSVD.py, dataset/transform.py) on a small synthetic caption table.  Build container only:

    python tests/golden/make_data_golden.py        ->  tests/golden/data_pipeline.json

The one reference step that cannot run on Python 3 is TrimExceptAscii (``str.decode``, dataset/transform.py:79-82, a Python-2
idiom); its effect -- dropping non-ASCII characters -- is applied to the captions before they reach the reference code.
Nothing is copied into the product; the script only *calls* the reference.
"""
import json
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import import_reference, REF          # stubs h5py / tensorboardX / coco_caption, puts the reference on sys.path

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data_pipeline.json")

ROWS = [  # VideoID, Start, End, Language, Description
    ("vidA", 0, 10, "English", "A man is playing a guitar."),
    ("vidA", 0, 10, "English", "a MAN plays the Guitar!"),
    ("vidA", 0, 10, "German", "Ein Mann spielt Gitarre"),
    ("vidB", 5, 9, "English", "Two dogs are running, fast; in the park"),
    ("vidB", 5, 9, "English", None),
    ("vidB", 5, 9, "English", "the dogs run in a park"),
    ("vidC", 1, 2, "English", "A woman is slicing an onion — café style"),
    ("vidC", 1, 2, "English", "someone is cooking"),
    ("vidD", 3, 8, "English", "a a a a a a a a a a a a long caption that goes on and on and on"),
    ("vidD", 3, 8, "English", "A man is cooking a fish."),
]
CAPTION_MAX_LEN, MIN_COUNT = 8, 2
PROBES = ["A man is playing a guitar.", "the dogs RUN!!", "zebra unknown words only", "", "a a a a a a a a a a a a a",
          "A woman is slicing an onion — café style", "man, cooking: a fish"]


def ascii_only(s):
    return s.encode("ascii", "ignore").decode("ascii") if isinstance(s, str) else s


def main():
    import_reference()
    import pandas as pd
    from dataset import MSVD as ref_msvd, transform as rt

    class Compose:              # torchvision.transforms.Compose (may be stubbed in this container)
        def __init__(self, fs): self.fs = fs
        def __call__(self, x):
            for f in self.fs:
                x = f(x)
            return x

    df = pd.DataFrame(ROWS, columns=["VideoID", "Start", "End", "Language", "Description"])
    csv_text = df.to_csv(index=False)
    ref_df = df.copy()
    ref_df["Description"] = ref_df["Description"].map(ascii_only)       # TrimExceptAscii, see the module docstring
    with tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False) as f:
        ref_df.to_csv(f, index=False)
        path = f.name
    sentence = Compose([rt.RemovePunctuation(), rt.Lowercase(), rt.SplitWithWhiteSpace(), rt.Truncate(CAPTION_MAX_LEN)])
    vocab = ref_msvd.MSVDVocab(path, {'<PAD>': 0, '<SOS>': 1, '<EOS>': 2}, MIN_COUNT, transform=sentence)
    caption = Compose([sentence, rt.ToIndex(vocab.word2idx), rt.PadLast(vocab.word2idx['<EOS>']),
                       rt.PadToLength(vocab.word2idx['<PAD>'], vocab.max_sentence_len + 1)])
    out = {"csv": csv_text, "caption_max_len": CAPTION_MAX_LEN, "min_count": MIN_COUNT,
           "word2idx": vocab.word2idx, "n_vocabs": vocab.n_vocabs, "n_words": vocab.n_words,
           "n_vocabs_untrimmed": vocab.n_vocabs_untrimmed, "n_words_untrimmed": vocab.n_words_untrimmed,
           "max_sentence_len": vocab.max_sentence_len,
           "probes": [{"text": p, "words": sentence(ascii_only(p)), "ids": [int(i) for i in caption(ascii_only(p))]} for p in PROBES]}
    # frame sampling (transform.py:9-62): frame k of an n-frame clip is the vector [k, k + 0.5]
    samp = []
    for method, cls in (("uniform", rt.UniformSample), ("random", rt.RandomSample), ("uniform_jitter", rt.UniformJitterSample)):
        for n in (5, 28, 29, 60, 333):
            frames = [np.array([k, k + 0.5], dtype=np.float32) for k in range(n)]
            np.random.seed(1000 + n)
            got = rt.ZeroPadIfLessThan(28)(list(cls(28)(frames)))
            samp.append({"method": method, "n": n, "seed": 1000 + n, "first_col": [float(x[0]) for x in got]})
    out["sampling"] = samp
    # dataset pairing + collate (MSVD.py:236-262, 53-76): clip ids and the order of the (clip, caption) pairs
    ds_caps = {}
    d2 = ref_df[ref_df["Language"] == "English"]
    d2 = d2[pd.notnull(d2["Description"])]
    for video_id, start, end, c in d2[["VideoID", "Start", "End", "Description"]].values:
        ds_caps.setdefault("{}_{}_{}".format(video_id, start, end), []).append(c)
    out["pairs"] = [[vid, c] for vid in ("vidB_5_9", "vidA_0_10", "vidD_3_8") for c in ds_caps[vid]]      # feature-file order B, A, D (no C)
    os.unlink(path)
    json.dump(out, open(OUT, "w"), indent=1, sort_keys=True)
    print("wrote", OUT, "vocab", vocab.n_vocabs, "max_sentence_len", vocab.max_sentence_len)


if __name__ == "__main__":
    if not os.path.isdir(REF):
        sys.exit("reference not mounted; golden files are committed, nothing to do")
    main()
