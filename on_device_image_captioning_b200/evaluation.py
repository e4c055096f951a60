"""The evaluation loop of the reference's test.py on the accelerated path.

``evaluate_model`` keeps the signature and the batching rules of reference test.py:141-275 up to the point where the
predictions are strings: sub-batches of ``parallel_batches`` items (the last one ragged), images stacked / features padded
exactly as the reference does, ``model(enc_x=..., enc_x_num_pads=..., mode="beam_search", **kwargs)`` per sub-batch,
``" ".join(words[1:-1])`` per caption, and the ``(pred_dict, gts_dict)`` return layout COCOEvalCap consumes
(test.py:230-249).  The COCO metrics themselves (PTB tokenizer, METEOR, SPICE: Java subprocesses under eval/) are not part
of the inference path; pass ``scorer`` to run them.

Differences, both optional: ``shard=True`` decodes rank r's sub-batches only and all-gathers the token ids (the reference
replicates the whole evaluation on every rank, test.py:310), and the per-caption Python work runs once per
sub-batch on one device->host copy (``batch_tokens2words``) instead of once per image.
"""
from __future__ import annotations

import math
from time import time
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

from .language_utils import convert_allsentences_idx2word


def compute_num_pads(list_bboxes) -> List[int]:
    """reference utils/language_utils.py:4-13: pads needed to bring every item to the longest one."""
    max_len = max((len(b) for b in list_bboxes), default=-1)
    return [max_len - len(b) for b in list_bboxes]


def sub_batch_ranges(num_samples: int, sb_size: int) -> List[Tuple[int, int]]:
    """The reference's sub-batch boundaries (test.py:165-175): ceil(n / sb) ranges, the last one takes the remainder."""
    n_it = math.ceil(num_samples / sb_size)
    return [(i * sb_size, num_samples if i == n_it - 1 else (i + 1) * sb_size) for i in range(n_it)]


def evaluate_model(ddp_model, y_idx2word_list, beam_size, max_seq_len, sos_idx, eos_idx, rank, ddp_sync_port=None,
                   parallel_batches=16, indexes=[0], data_loader=None, dataset_split=None, use_images_instead_of_features=False,
                   verbose=True, stanford_model_path=None, scorer: Optional[Callable[[Dict, Dict], Dict]] = None, shard: bool = False):
    """reference test.py:141-275.  ``data_loader`` needs get_images_by_idx / get_bboxes_by_idx / get_captions_by_idx
    (data/coco_dataloader.py); ``scorer(gts_dict, pred_dict)`` replaces the COCOEvalCap call when given.
    Returns (pred_dict, gts_dict) on rank 0 (on every rank when not distributed), (None, None) elsewhere."""
    start_time = time()
    import torch.distributed as dist
    world = dist.get_world_size() if (shard and dist.is_available() and dist.is_initialized()) else 1
    my_rank = dist.get_rank() if world > 1 else 0
    num_samples = len(indexes)
    ranges = sub_batch_ranges(num_samples, parallel_batches)
    predictions: Dict[int, str] = {}
    validate_y: Dict[int, Sequence[str]] = {}
    kwargs = {"beam_size": beam_size, "beam_max_seq_len": max_seq_len, "sample_or_max": "max", "how_many_outputs": 1,
              "sos_idx": sos_idx, "eos_idx": eos_idx}
    was_training = getattr(ddp_model, "training", False)
    ddp_model.eval()
    local: List[Tuple[int, List[List[int]]]] = []
    with torch.no_grad():
        for sb_it, (from_idx, to_idx) in enumerate(ranges):
            ids = list(range(from_idx, to_idx))
            for i in ids:
                validate_y[i] = data_loader.get_captions_by_idx(i, dataset_split=dataset_split)
            if world > 1 and sb_it % world != my_rank:
                continue
            if use_images_instead_of_features:
                x = torch.cat([data_loader.get_images_by_idx(i, dataset_split=dataset_split).unsqueeze(0) for i in ids]).to(rank)
                pads = [0] * x.size(0)
            else:
                x = [data_loader.get_bboxes_by_idx(i, dataset_split=dataset_split) for i in ids]
                x = torch.nn.utils.rnn.pad_sequence(x, batch_first=True).to(rank)
                pads = compute_num_pads(x)                  # on the padded tensor, as the reference does (test.py:194)
            output_words, _ = ddp_model(enc_x=x, enc_x_num_pads=pads, mode="beam_search", **kwargs)
            local.append((from_idx, [output_words[i][0] for i in range(len(output_words))]))
    if was_training:
        ddp_model.train()
    if world > 1:
        gathered: List = [None] * world
        dist.all_gather_object(gathered, local)
        local = [item for part in gathered for item in part]
    for from_idx, toks in local:
        for k, sentence in enumerate(convert_allsentences_idx2word(toks, y_idx2word_list)):
            predictions[from_idx + k] = " ".join(sentence[1:-1])      # remove EOS and SOS (test.py:221-224)
    if my_rank != 0 and world > 1:
        return None, None
    gts_dict = {i: [{"image_id": i, "caption": c} for c in validate_y[i]] for i in range(num_samples)}
    pred_dict = {i: [{"image_id": i, "caption": predictions[i]}] for i in range(num_samples)}
    if scorer is not None and verbose:
        score_results = scorer(gts_dict, pred_dict)
        elapsed = time() - start_time
        print("Evaluation Phase over " + str(num_samples) + " BeamSize: " + str(beam_size) + "  elapsed: " +
              str(int(elapsed / 60)) + " m " + str(int(elapsed % 60)) + " s")
        print(score_results)
    return pred_dict, gts_dict
