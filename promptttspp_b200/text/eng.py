"""Phoneme symbol table and id conversion -- the host-side string work in front of the hot path
(reference: promptttspp/text/eng.py:9-158; callers app.py:60-62, egs/proposed/bin/synthesize.py:165).

The inventory is CMUdict's ARPAbet as g2p_en emits it: every vowel bare and with the stress digits 0-2, the consonants,
sorted, followed by the three silence/unknown marks; ids 0/1/2 are PAD/BOS/EOS (num_vocab = 90, the embedding table size
of the shipped configs).  `batch_text_to_sequence` is the batched form the serving front end uses: one pass over all
requests into a pinned, padded id matrix ready for a single H2D copy.
"""
from typing import List, Sequence

import torch

PAD = "_"
BOS = "^"
EOS = "$"

_VOWELS = ["AA", "AE", "AH", "AO", "AW", "AY", "EH", "ER", "EY", "IH", "IY", "OW", "OY", "UH", "UW"]
_CONSONANTS = ["B", "CH", "D", "DH", "F", "G", "HH", "JH", "K", "L", "M", "N", "NG", "P", "R", "S", "SH", "T", "TH", "V",
               "W", "Y", "Z", "ZH"]
phonemes = sorted([v + s for v in _VOWELS for s in ("", "0", "1", "2")] + _CONSONANTS) + ["spn", "sil", "sp"]
symbols = [PAD, BOS, EOS] + phonemes
symbol2id = {s: i for i, s in enumerate(symbols)}


def symbol_to_id(symbol):
    return symbol2id[symbol]


def id_to_symbol(idnum):
    return symbols[idnum]


def num_vocab():
    return len(symbols)


def text_to_sequence(text, add_special_token=True):
    """Space-separated phoneme string -> list of ids, framed by BOS/EOS (eng.py:117-140).  Unknown symbols raise
    KeyError like the reference."""
    seq = [symbol2id[ph] for ph in text.split()]
    if add_special_token:
        seq = [symbol2id[BOS]] + seq + [symbol2id[EOS]]
    return seq


def sequence_to_text(seq, remove_special_token=False):
    """ids -> list of phoneme symbols (eng.py:143-158)."""
    seq = list(seq)
    if remove_special_token:
        seq = seq[1:-1]
    return [symbols[int(s)] for s in seq]


def batch_text_to_sequence(texts: Sequence[str], add_special_token=True, pin=True):
    """N phoneme strings -> (ids [N, Lmax] int64 padded with PAD, lengths [N] int64), pinned for one async H2D copy."""
    seqs: List[List[int]] = [text_to_sequence(t, add_special_token) for t in texts]
    lens = torch.tensor([len(s) for s in seqs], dtype=torch.int64)
    out = torch.zeros(len(seqs), int(lens.max()) if seqs else 0, dtype=torch.int64)
    for i, s in enumerate(seqs):
        out[i, : len(s)] = torch.tensor(s, dtype=torch.int64)
    if pin and torch.cuda.is_available():
        out, lens = out.pin_memory(), lens.pin_memory()
    return out, lens
