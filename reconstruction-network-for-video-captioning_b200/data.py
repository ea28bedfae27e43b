"""Synthetic MSVD-shaped batches.  The reference's dataset/ loader (HDF5 features + CSV
captions) is out of scope; this produces tensors of exactly the shapes and value conventions it would:
feats (B, T, E) fp32; targets (caption_max_len + 1, B) int64 = word ids, then <EOS>=2, then <PAD>=0 (no <SOS>
inside targets, dataset/MSVD.py:111-117); masks = targets > 0 (train.py:246)."""
import torch

PAD, SOS, EOS = 0, 1, 2


def synthetic_batch(B, T, E, V, caption_max_len=30, seed=1234, full_length_first=True):
    g = torch.Generator().manual_seed(seed)
    feats = torch.randn(B, T, E, generator=g, dtype=torch.float32)
    lens = torch.randint(min(3, caption_max_len), caption_max_len + 1, (B,), generator=g)
    if full_length_first:
        lens[0] = caption_max_len          # => L = caption_max_len + 1 decoded steps every iteration (fixed work)
    targets = torch.zeros(caption_max_len + 1, B, dtype=torch.long)
    for b in range(B):
        n = int(lens[b])
        targets[:n, b] = torch.randint(3, V, (n,), generator=g)
        targets[n, b] = EOS
    return feats, targets, targets > PAD


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of n_items for `rank` of `world` (remainder spread over the first ranks)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)
