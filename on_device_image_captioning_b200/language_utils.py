"""Mirror of the detokenisation helpers of the reference's utils/language_utils.py:60-93 (same names, same results),
plus the batched form the accelerated path returns its captions in.  Host-side string work: there is nothing to
accelerate, it only has to keep the reference's exact rules (skip SOS, stop at EOS, final full stop, capitalise)."""
from __future__ import annotations

from typing import List, Sequence


def convert_vector_word2idx(sentence, word2idx_dict):
    return [word2idx_dict[word] for word in sentence]


def convert_allsentences_word2idx(sentences, word2idx_dict):
    return [convert_vector_word2idx(s, word2idx_dict) for s in sentences]


def convert_vector_idx2word(sentence, idx2word_list):
    return [idx2word_list[idx] for idx in sentence]


def convert_allsentences_idx2word(sentences, idx2word_list):
    return [convert_vector_idx2word(s, idx2word_list) for s in sentences]


def tokens2description(tokens, idx2word_list, sos_idx, eos_idx):
    """reference utils/language_utils.py:82-93 (raises IndexError on a caption without words, like the reference)."""
    desc = []
    for tok in tokens:
        if tok == sos_idx:
            continue
        if tok == eos_idx:
            break
        desc.append(tok)
    desc = convert_vector_idx2word(desc, idx2word_list)
    desc[-1] = desc[-1] + "."
    pred = " ".join(desc).capitalize()
    return pred


def batch_tokens2description(tokens, lengths, idx2word_list, sos_idx, eos_idx) -> List[str]:
    """Device results of beam_search -- tokens (B, how_many, L) int32 padded with -1, lengths (B, how_many) -- to the
    best caption string of every image (one device->host copy for the whole batch)."""
    tok = tokens.cpu().tolist() if hasattr(tokens, "cpu") else tokens
    ln = lengths.cpu().tolist() if hasattr(lengths, "cpu") else lengths
    return [tokens2description(tok[b][0][: ln[b][0]], idx2word_list, sos_idx, eos_idx) for b in range(len(tok))]
