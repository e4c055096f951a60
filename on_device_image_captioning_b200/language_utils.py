"""Detokenisation helpers with the names and results of the reference's utils/language_utils.py:60-93, plus the batched
form the accelerated path returns its captions in.  Host-side string work -- nothing to accelerate; it only has to keep
the reference's rules: SOS tokens are dropped wherever they occur, the caption ends at the first EOS, the last word gets
a full stop, and the sentence is capitalised with str.capitalize() (which also lower-cases the rest)."""
from __future__ import annotations

from itertools import takewhile
from typing import Dict, List, Sequence


def convert_vector_word2idx(sentence: Sequence[str], word2idx_dict: Dict[str, int]) -> List[int]:
    return list(map(word2idx_dict.__getitem__, sentence))


def convert_allsentences_word2idx(sentences, word2idx_dict):
    return [convert_vector_word2idx(one, word2idx_dict) for one in sentences]


def convert_vector_idx2word(sentence: Sequence[int], idx2word_list: Sequence[str]) -> List[str]:
    return list(map(idx2word_list.__getitem__, sentence))


def convert_allsentences_idx2word(sentences, idx2word_list):
    return [convert_vector_idx2word(one, idx2word_list) for one in sentences]


def tokens2description(tokens, idx2word_list, sos_idx, eos_idx) -> str:
    """Token ids -> caption string.  A caption without any word raises IndexError, as the reference does (its
    ``desc[-1]`` on an empty list)."""
    body = [t for t in takewhile(lambda t: t != eos_idx, tokens) if t != sos_idx]
    words = convert_vector_idx2word(body, idx2word_list)
    if not words:
        raise IndexError("list index out of range")
    words[-1] += "."
    return " ".join(words).capitalize()


def batch_tokens2description(tokens, lengths, idx2word_list, sos_idx, eos_idx) -> List[str]:
    """Device results of beam_search -- tokens (B, how_many, L) int32 padded with -1, lengths (B, how_many) -- to the
    best caption string of every image (one device->host copy for the whole batch)."""
    tok = tokens.cpu().tolist() if hasattr(tokens, "cpu") else tokens
    ln = lengths.cpu().tolist() if hasattr(lengths, "cpu") else lengths
    return [tokens2description(row[0][: n[0]], idx2word_list, sos_idx, eos_idx) for row, n in zip(tok, ln)]
