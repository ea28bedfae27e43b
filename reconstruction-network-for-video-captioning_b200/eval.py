"""Inference mirrors of the reference's eval.greedy_search (eval.py:19-33)."""
from __future__ import annotations

import torch


def greedy_search(config, decoder, input, hidden, encoder_outputs):
    """Same signature and return value as eval.greedy_search: a list (one entry per generated step) of lists of
    B token ids.  ``input`` must be the <SOS> row and ``hidden`` the zero state (what eval.evaluate passes,
    eval.py:130-140); the argmax feedback loop runs entirely on the device (one host read at the end instead of
    B per step, eval.py:25)."""
    if not bool((input == 1).all()):
        raise NotImplementedError("greedy_search: the device loop starts from <SOS>; use Decoder.forward for custom starts")
    ids, n = decoder.greedy(encoder_outputs, config.caption_max_len + 1)
    n = int(n.item())
    return ids[:n].tolist()
