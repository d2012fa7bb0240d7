"""CPU: the callers behind the decode path (evaluate.py:15-74, train.py:151-166): ids -> words, NLTK-compatible BLEU with
SmoothingFunction().method4, sentence statistics, and the synthetic DataReader shim."""
import math

import numpy as np
import pytest

import myimagecaptioningmodel_b200 as M
from myimagecaptioningmodel_b200 import evaluation as EV

HYP1 = "It is a guide to action which ensures that the military always obeys the commands of the party".split()
REF1 = "It is a guide to action that ensures that the military will forever heed Party commands".split()
REF2 = "It is the guiding principle which guarantees the military forces always being under the command of the Party".split()
REF3 = "It is the practical guide for the army always to heed the directions of the party".split()


def test_sentence_bleu_matches_the_nltk_documentation_values():
    """golden values from the doctests of nltk.translate.bleu_score (sentence_bleu / modified_precision)"""
    assert abs(EV.sentence_bleu([REF1, REF2, REF3], HYP1, smoothing=None) - 0.5045666840058485) < 1e-12
    assert round(EV.sentence_bleu([REF1, REF2, REF3], HYP1, weights=(1. / 5,) * 5, smoothing=None), 4) == 0.3920
    # modified_precision doctest: 'the' x7 against two references -> 2/7
    hyp = "the the the the the the the".split()
    assert EV.modified_precision(["the cat is on the mat".split(), "there is a cat on the mat".split()], hyp, 1) == (2, 7)
    assert EV.modified_precision([REF1, REF2, REF3], HYP1, 1) == (17, 18)
    assert EV.modified_precision([REF1, REF2, REF3], HYP1, 2) == (10, 17)
    # brevity penalty doctest: hypothesis of 12 words against references of 13 / 2 words picks 13 -> exp(1 - 13/12)
    assert EV.closest_ref_length([["a"] * 13, ["a"] * 2], 12) == 13
    assert abs(EV.brevity_penalty(13, 12) - math.exp(1 - 13 / 12)) < 1e-15
    assert EV.closest_ref_length([["a"] * 11, ["a"] * 13], 12) == 11       # tie -> the shorter reference
    # all n-gram orders matched: smoothing does nothing
    assert abs(EV.sentence_bleu([REF1, REF2, REF3], HYP1) - 0.5045666840058485) < 1e-12
    assert EV.sentence_bleu([HYP1], HYP1) == 1.0
    assert EV.sentence_bleu([REF1], "completely unrelated words here".split()) == 0


def test_method4_smoothing():
    hyp, ref = "a b c x".split(), "a b c d".split()         # 1-gram 3/4, 2-gram 2/3, 3-gram 1/2, 4-gram 0/1
    got = EV.sentence_bleu([ref], hyp)
    p4 = (3 + 5.0 / math.log(4)) / 1                        # NLTK 3.5: (i + k / ln(len)) / denominator with i = 3
    want = math.exp(0.25 * (math.log(3 / 4) + math.log(2 / 3) + math.log(1 / 2) + math.log(p4)))
    assert abs(got - want) < 1e-12
    later = EV.sentence_bleu([ref], hyp, method4_version="3.6")
    p4b = 1 / (2 * 5.0 / math.log(4))
    assert abs(later - math.exp(0.25 * (math.log(3 / 4) + math.log(2 / 3) + math.log(1 / 2) + math.log(p4b)))) < 1e-12
    with pytest.raises(ZeroDivisionError):                  # the case evaluate.py:36-37 guards against
        EV.sentence_bleu([["a", "b"]], ["a"])
    # unigram-only weights as in evaluate.py:56 ([1,0,0,0]): higher orders do not matter
    assert abs(EV.sentence_bleu([ref], hyp, weights=[1, 0, 0, 0]) - 0.75) < 1e-12


def test_filter_and_calc_bleu():
    iw = {0: "<pad>", 1: "<unk>", 2: "<start>", 3: "<stop>", 4: "a", 5: "b", 6: "c", 7: "d", 8: "x"}
    assert EV.filter([4, 0, 5, 3, 6], iw) == ["a", "b"]                    # evaluate.py:15-25
    assert EV.filter([3, 4], iw) == [] and EV.words2sentence(["a", "b"]) == "a b"
    assert M.filter_caption([4.0, 0.0, 5.0, 3.0, 6.0]) == [4, 5]           # the id-level twin used by the decoder tests
    pred = np.array([[4, 5, 6, 8, 3, 0], [4, 3, 0, 0, 0, 0], [4, 5, 6, 7, 3, 0]], dtype="float32")   # model:122 float ids
    real = [[["a", "b", "c", "d"]], [["a", "b"]], [["a", "b", "c", "d"]]]
    got = EV.calc_bleu(pred, real, iw)
    want = (EV.sentence_bleu(real[0], ["a", "b", "c", "x"]) + 0 + 1.0) / 3   # one-word caption: skipped, still counted
    assert abs(got - want) < 1e-12
    assert EV.calc_bleu(pred.astype("int64").tolist(), real, iw) == got


def test_evaluate_loop_and_synthetic_reader():
    cfg = M.DecoderConfig(hidden_dim=32, embedding_size=16, vocab_size=53, num_positions=7, sentence_length=9,
                          infer_max_length=9)
    rd = EV.SyntheticReader(cfg, n_samples=10, encoder_channel=16, seed=1, refs_per_image=2)
    with pytest.raises(ValueError):
        rd.get_reader(4, "val")                                            # reader.py:52-53
    train = list(rd.get_reader(4, "train")())
    assert [len(b) for b in train] == [4, 4, 2]
    img, cap = train[0][0]
    assert img.shape == (16, 7) and img.dtype == np.float32 and cap.dtype == np.int64 and cap[0] == cfg.start_idx
    test = rd.get_reader(4, "test")
    refs0 = test().__next__()[0][1]
    assert len(refs0) == 2 and all(w.startswith("w") for w in refs0[0])

    def perfect_decoder(images):            # "decodes" every image to its first reference
        out = np.zeros((len(images), cfg.infer_max_length), dtype="float32")
        for j, im in enumerate(images):
            i = next(i for i in range(rd.n) if np.array_equal(rd._maps[i], im))
            c = rd._caps[i, 0][1:]
            out[j, :len(c)] = c
        return out

    bleu, distinct = EV.evaluate(perfect_decoder, test(), rd.index_word)
    assert all(abs(b - 1.0) < 1e-12 for b in bleu) and distinct == 10      # evaluate.py:69-74
    bleu0, distinct0 = EV.evaluate(lambda im: np.full((len(im), 9), 5.0, dtype="float32"), test(), rd.index_word)
    assert distinct0 == 1 and bleu0[4] < 0.2
    with pytest.raises(ValueError):
        EV.evaluate(perfect_decoder, [], rd.index_word)


def test_caption_images_with_a_stub_engine():
    """infer.py:26-38 after the encoder: float32 ids in, sentences out (the engine is stubbed; the GPU version of this call is
    exercised in tests/test_train_loop_gpu.py)"""
    import torch
    import types
    cfg = M.DecoderConfig(hidden_dim=32, embedding_size=16, vocab_size=53, num_positions=7, sentence_length=9,
                          infer_max_length=9)
    rd = EV.SyntheticReader(cfg, n_samples=4, encoder_channel=16, seed=1, refs_per_image=1)
    ids = torch.tensor([[5., 6., 7., 3., 9., 0., 0., 0., 0.], [3., 0., 0., 0., 0., 0., 0., 0., 0.]])
    eng = types.SimpleNamespace(cfg=cfg, decode_map=lambda x, beam, finish_on_stop: (None, None, ids))
    out_ids, sentences = EV.caption_images(eng, None, rd.index_word)
    assert out_ids.dtype == np.float32 and sentences == ["w5 w6 w7", ""]
