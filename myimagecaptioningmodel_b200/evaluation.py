"""The callers behind the decode path: ids -> words, BLEU and sentence statistics.

Mirrors /root/reference/ImageCaptioning/evaluate.py:15-74 and the dev loop of train.py:151-166.  Pure host code: the ids
come from capdec_decode (float32 [B,35], the reference's return type), everything here runs on the CPU exactly as it
does in the reference.

`sentence_bleu` restates nltk.translate.bleu_score (evaluate.py:3-4 imports it; NLTK is not installed here and the
reference pins no version).  The arithmetic follows NLTK 3.4 / 3.5 -- the releases current when the reference was
written (mid 2020), and the only ones whose `method4` raises ZeroDivisionError for one-word hypotheses, which is what the
`len(p) <= 1` guard of evaluate.py:36-37 protects against:
  modified n-gram precision clipped over all references, closest reference length, brevity penalty,
  SmoothingFunction().method4:  p_n with a zero numerator becomes (i + k / ln(len(hyp))) / denominator, k = 5, i = 0-based
  order index.  `method4_version="3.6"` selects the later form 1 / (2^j * k / ln(len)) / denominator instead.
"""
from __future__ import annotations

import math
from collections import Counter
from fractions import Fraction
from typing import Callable, Iterable, List, Optional, Sequence

import numpy as np


def filter(p, index_word, stop_tag: int = 3, padding_tag: int = 0):   # noqa: A001  (the reference's name, evaluate.py:15)
    """evaluate.py:15-25: ids -> words.  Everything from the first <stop> on is dropped, <pad> is skipped."""
    ids = list(p)
    if stop_tag in ids:
        ids = ids[:ids.index(stop_tag)]
    return [index_word[i] for i in ids if i != padding_tag]


def words2sentence(words):
    """evaluate.py:43-44"""
    return " ".join(words)


# ----------------------------------------------------------------------------------------------------------------
# nltk.translate.bleu_score.sentence_bleu + SmoothingFunction().method4
# ----------------------------------------------------------------------------------------------------------------
def _ngrams(seq, n):
    return [tuple(seq[i:i + n]) for i in range(len(seq) - n + 1)]


def modified_precision(references, hypothesis, n):
    counts = Counter(_ngrams(hypothesis, n)) if len(hypothesis) >= n else Counter()
    max_counts = {}
    for ref in references:
        ref_counts = Counter(_ngrams(ref, n)) if len(ref) >= n else Counter()
        for ng in counts:
            max_counts[ng] = max(max_counts.get(ng, 0), ref_counts[ng])
    clipped = {ng: min(c, max_counts[ng]) for ng, c in counts.items()}
    return sum(clipped.values()), max(1, sum(counts.values()))


def closest_ref_length(references, hyp_len):
    return min((len(r) for r in references), key=lambda ref_len: (abs(ref_len - hyp_len), ref_len))


def brevity_penalty(closest_ref_len, hyp_len):
    if hyp_len > closest_ref_len:
        return 1.0
    if hyp_len == 0:
        return 0.0
    return math.exp(1 - closest_ref_len / hyp_len)


def method4(p_n, hyp_len, k: float = 5.0, version: str = "3.5"):
    """SmoothingFunction().method4 on a list of (numerator, denominator) pairs -> list of floats / Fractions."""
    out = []
    incvnt = 1
    for i, (num, den) in enumerate(p_n):
        if num == 0 and hyp_len != 0 and version == "3.5":
            out.append((i + 1 * k / math.log(hyp_len)) / den)          # ZeroDivisionError for hyp_len == 1, as in NLTK 3.5
        elif num == 0 and hyp_len > 1 and version != "3.5":
            out.append(1 / (2 ** incvnt * k / math.log(hyp_len)) / den)
            incvnt += 1
        else:
            out.append(Fraction(num, den))
    return out


def sentence_bleu(references, hypothesis, weights=(0.25, 0.25, 0.25, 0.25), smoothing: Optional[str] = "method4",
                  method4_version: str = "3.5"):
    """nltk sentence_bleu(references, hypothesis, weights, smoothing_function) for smoothing in {None (method0), 'method4'}."""
    p_n = [modified_precision(references, hypothesis, i) for i, _ in enumerate(weights, start=1)]
    hyp_len = len(hypothesis)
    bp = brevity_penalty(closest_ref_length(references, hyp_len), hyp_len)
    if p_n[0][0] == 0:
        return 0                                                        # no unigram match
    if smoothing == "method4":
        vals = method4(p_n, hyp_len, version=method4_version)
    elif smoothing is None:
        # SmoothingFunction().method0: a zero precision is replaced by sys.float_info.min (and a warning is printed)
        vals = [Fraction(n, d) if n != 0 else 2.2250738585072014e-308 for n, d in p_n]
    else:
        raise ValueError("unsupported smoothing {}".format(smoothing))
    s = math.fsum(w * math.log(v) for w, v in zip(weights, vals))
    return bp * math.exp(s)


def _as_id_lists(pred):
    """The decoder returns float32 ids (model:122): round to int like evaluate.py:30-32."""
    if isinstance(pred, np.ndarray):
        if pred.dtype == "float32":
            pred = np.rint(pred).astype("int32")
        return pred.tolist()
    return pred


def calc_bleu(pred, real, index_word, weights=(0.25, 0.25, 0.25, 0.25), stop_tag: int = 3, padding_tag: int = 0,
              method4_version: str = "3.5"):
    """evaluate.py:28-40: mean over the batch of the smoothed sentence BLEU.  Captions of <= 1 word score 0 but still
    count in the denominator (the guard that keeps NLTK 3.5's method4 away from ln(1) = 0)."""
    pred = _as_id_lists(pred)
    captions = [filter(p, index_word, stop_tag, padding_tag) for p in pred]
    scores = [sentence_bleu(refs, words, weights=weights, smoothing="method4", method4_version=method4_version)
              for words, refs in zip(captions, real) if len(words) > 1]
    return sum(scores) / len(pred)


def evaluate(decode_fn: Callable, batches: Iterable, index_word, sentence_statistics: bool = True, stop_tag: int = 3,
             padding_tag: int = 0):
    """evaluate.py:46-74 with the Paddle inference program replaced by `decode_fn(images) -> float32 ids [B,T]`
    (e.g. ``lambda x: engine.decode_map(x)[2].cpu().numpy()``).  `batches` yields lists of (image, references) like
    DataReader.get_reader(batch_size, 'test') (reader.py:65-72).  Returns ([BLEU-1, BLEU-2, BLEU-3, BLEU-4, their mean],
    number of distinct sentences): per-order scores use one-hot weights and are averaged over BATCHES, not samples, exactly
    like the reference (a short last batch weighs as much as a full one)."""
    one_hot = np.eye(4).tolist()
    per_batch = []                       # rows: [BLEU-1..4] of one batch
    said = set()
    for data in batches:
        images = np.stack([d[0] for d in data])
        refs = [d[1] for d in data]
        ids = np.asarray(decode_fn(images))
        per_batch.append([calc_bleu(ids, refs, index_word, w, stop_tag, padding_tag) for w in one_hot])
        if sentence_statistics:
            said.update(words2sentence(filter(p, index_word, stop_tag, padding_tag)) for p in _as_id_lists(ids))
    if not per_batch:
        raise ValueError("no batches to evaluate")
    orders = np.mean(np.asarray(per_batch, dtype=np.float64), axis=0).tolist()
    return orders + [sum(orders) / 4], len(said)


def caption_images(engine, encoder_map, index_word, beam: int = 1, finish_on_stop: bool = False):
    """infer.py:26-38 from the encoder output on: decode and turn the ids into sentences.  beam = 1 without
    finished-handling is the reference's greedy `eval_network` (model:185-189); a wider beam returns the best hypothesis.
    Returns (float32 ids [B,T] as the reference prints them, list of sentences)."""
    cfg = engine.cfg
    ids_f32 = engine.decode_map(encoder_map, beam=beam, finish_on_stop=finish_on_stop)[2].cpu().numpy()
    sentences = [words2sentence(filter(p, index_word, cfg.stop_idx, cfg.padding_idx)) for p in _as_id_lists(ids_f32)]
    return ids_f32, sentences


class SyntheticReader:
    """Stand-in for reader.DataReader (reader.py:31-94) on synthetic data: `get_reader(batch_size, mode)` returns a
    callable yielding batches (lists of (encoder_map [X,k] float32, caption)) where caption is an int64 id array in
    'train' mode and a list of reference word lists in 'dev' / 'test' mode; `index_word` / `word_index` are the
    dictionaries.  Vocabulary ids follow the tokenizer: <pad>=0 <unk>=1 <start>=2 <stop>=3."""

    def __init__(self, cfg, n_samples: int = 64, encoder_channel: int = 1280, seed: int = 0, refs_per_image: int = 5):
        from . import synth
        self.cfg, self.n = cfg, n_samples
        self._maps = synth.make_encoder_map(cfg, n_samples, encoder_channel, seed).numpy()
        self._caps = synth.make_captions(cfg, n_samples * refs_per_image, seed).numpy().reshape(
            n_samples, refs_per_image, cfg.sentence_length)
        self.index_word = {i: w for i, w in enumerate(["<pad>", "<unk>", "<start>", "<stop>"])}
        for i in range(4, cfg.vocab_size):
            self.index_word[i] = "w%d" % i
        self.word_index = {w: i for i, w in self.index_word.items()}

    def _refs(self, i):
        return [filter(c[1:].tolist(), self.index_word, self.cfg.stop_idx, self.cfg.padding_idx) for c in self._caps[i]]

    def get_reader(self, batch_size: Optional[int] = None, mode: str = "train"):
        if mode not in ["train", "dev", "test"]:
            raise ValueError("DataReader不支持 {} 模式".format(mode))                    # reader.py:52-53

        def samples():
            for i in range(self.n):
                yield (self._maps[i], self._caps[i, 0].astype("int64")) if mode == "train" else (self._maps[i], self._refs(i))

        def batched():
            buf: List = []
            for s in samples():
                buf.append(s)
                if len(buf) == batch_size:
                    yield buf
                    buf = []
            if buf:
                yield buf                                                                # fluid.io.batch, drop_last=False
        return samples if batch_size is None else batched
