"""CPU tests of the boundary: the C-ABI library builds, loads and exports every symbol include/capdec.h declares;
host-side logic (config, parameter layout, sharding, error behaviour) works without a GPU; the product path fails
loudly when no GPU is present (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

import myimagecaptioningmodel_b200 as M
from myimagecaptioningmodel_b200 import _build, _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "capdec.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(capdec_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    path = _build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), "libcapdec.so does not export %s" % name
    assert sorted(_capi.EXPORTS) == declared            # the ctypes binding covers the whole header
    assert _capi.load_library().capdec_version() >= 100


def test_struct_layouts_match_header():
    text = open(os.path.join(ROOT, "include", "capdec.h")).read()
    cfg_body = re.search(r"typedef struct capdec_config \{(.*?)\} capdec_config;", text, re.S).group(1)
    cfg_body = re.sub(r"/\*.*?\*/", "", cfg_body, flags=re.S)
    fields = re.findall(r"int32_t\s+(\w+);", cfg_body)
    assert fields == [f[0] for f in _capi.CapdecConfig._fields_]
    prm_body = re.search(r"typedef struct capdec_params \{(.*?)\} capdec_params;", text, re.S).group(1)
    prm_body = re.sub(r"/\*.*?\*/", "", prm_body, flags=re.S)
    names = re.findall(r"const float\*\s+(\w+);", prm_body)
    assert names == _capi.PARAM_NAMES
    grd_body = re.search(r"typedef struct capdec_grads \{(.*?)\} capdec_grads;", text, re.S).group(1)
    assert re.findall(r"float\*\s+(\w+);", grd_body) == _capi.PARAM_NAMES
    assert ctypes.sizeof(_capi.CapdecConfig) == 48


def test_sm100_sass_is_present():
    """the built library really contains tcgen05 / TMA code for sm_100a (not a recompiled legacy path)"""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", _build.build()], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    # tcgen05.mma (single CTA and cta_group::2), TMA tensor loads / stores / reduce-adds (+ the 2-CTA and multicast forms),
    # TMA bulk copies, tcgen05.ld
    for mnemonic in ("UTCHMMA", "UTCHMMA.2CTA", "UTMALDG.3D", "UTMALDG.3D.2CTA", "UTMASTG.3D", "UTMAREDG.3D.ADD",
                     "UTCBAR.2CTA.MULTICAST", "UBLKCP", "LDTM"):
        assert mnemonic in sass, mnemonic
    # no GEMM runs on the legacy mma.sync path: the only HMMA in the library is the matrix-VECTOR product inside the
    # attention score kernel (tanh results as the A fragment, w_alpha as one B column; attention.cu)
    func = None
    hmma_funcs = set()
    for line in sass.splitlines():
        if "Function :" in line:
            func = line.split("Function :")[1].strip()
        elif "HMMA." in line and "UTCHMMA" not in line:
            hmma_funcs.add(func)
    assert hmma_funcs and all("attn_scores_kernel" in f for f in hmma_funcs), hmma_funcs


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    with pytest.raises(M.CapdecError):
        M.CaptionDecoderEngine()
    lib = M.load_library()
    h = ctypes.c_void_p()
    cfg = M.DecoderConfig().to_c()
    assert lib.capdec_create(ctypes.byref(cfg), 0, ctypes.byref(h)) != 0
    assert b"no CPU fallback" in lib.capdec_last_error() or b"CUDA" in lib.capdec_last_error()


def test_create_rejects_bad_arguments():
    lib = M.load_library()
    h = ctypes.c_void_p()
    assert lib.capdec_create(None, 0, ctypes.byref(h)) != 0
    bad = M.DecoderConfig(hidden_dim=1023).to_c()
    assert lib.capdec_create(ctypes.byref(bad), 0, ctypes.byref(h)) != 0
    assert b"multiples of 8" in lib.capdec_last_error()
    bad = M.DecoderConfig(num_positions=64).to_c()
    assert lib.capdec_create(ctypes.byref(bad), 0, ctypes.byref(h)) != 0
    assert lib.capdec_destroy(None) == 0
    assert lib.capdec_workspace_bytes(None) == 0


def test_config_and_parameter_layout():
    cfg = M.DecoderConfig()
    assert cfg.train_steps == 34                       # model:66
    shapes = M.param_shapes(cfg)
    assert list(shapes) == M.PARAM_NAMES
    assert shapes["lstm_w"] == (2304, 4096) and shapes["word_embedding"] == (12295, 256)
    n = sum(int(torch.tensor(s).prod()) for s in shapes.values())
    assert n == 21523464                               # decoder parameter count
    with pytest.raises(ValueError):
        M.DecoderConfig(dtype="fp16").to_c()
    with pytest.raises(ValueError):
        M.DecoderConfig(softmax_axis="rows").to_c()
    store = M.ParameterStore(M.DecoderConfig(hidden_dim=64, embedding_size=32, vocab_size=101, num_positions=5,
                                             sentence_length=8, infer_max_length=9), "cpu").init_reference(3, 0.1)
    assert store.num_parameters == sum(v.numel() for v in store.params.values())
    for name, t in store.params.items():
        # vector-load alignment of every view relative to the (512-byte aligned on CUDA) flat buffer
        assert (t.data_ptr() - store.flat.data_ptr()) % 256 == 0, name
    assert float(store.params["word_embedding"].abs().max()) <= 1.0
    sd = store.state_dict()
    store2 = M.ParameterStore(store.cfg, "cpu").load(sd)
    assert torch.equal(store2.flat, store.flat)
    with pytest.raises(KeyError):
        store2.load({})


def test_decoder_rejects_unknown_mode():
    with pytest.raises(ValueError):
        M.Decoder(1024, mode="test", engine=object())    # model:43-44


def test_shard_range():
    for n, w in [(4096, 8), (10, 3), (2, 4), (0, 2)]:
        spans = [M.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        M.shard_range(4, 4, 4)


def test_filter_caption():
    assert M.filter_caption([5.0, 0.0, 7.0, 3.0, 9.0]) == [5, 7]     # evaluate.py:15-25
    assert M.filter_caption([3.0, 4.0]) == []


def test_header_is_plain_c_and_links_against_the_library(tmp_path):
    """include/capdec.h is a C header (no C++ / torch types): a C99 program that uses every struct and takes the address of
    every entry point compiles with gcc and links against libcapdec.so; without a GPU capdec_create fails with a message."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    syms = _declared_symbols()
    src = ['#include "capdec.h"', "#include <stdio.h>", "typedef void (*fn_t)(void);", "int main(void) {",
           "  capdec_config cfg = {1024, 256, 12295, 49, 35, 35, 0, 2, 3, CAPDEC_DTYPE_BF16, CAPDEC_SOFTMAX_POSITION, 0};",
           "  capdec_params p = {0}; capdec_grads g = {0}; capdec_adaptor_params ap = {0}; capdec_adaptor_grads ag = {0};",
           "  capdec_handle h = 0; (void)p; (void)g; (void)ap; (void)ag;",
           "  fn_t fns[] = {" + ", ".join("(fn_t)%s" % s for s in syms) + "};",
           "  int rc = capdec_create(&cfg, 0, &h);",
           '  printf("%d %d %s\\n", (int)(sizeof(fns) / sizeof(fns[0])), rc, rc ? capdec_last_error() : "ok");',
           "  if (!rc) capdec_destroy(h);", "  return 0;", "}"]
    c_file = tmp_path / "use_capdec.c"
    c_file.write_text("\n".join(src))
    exe = str(tmp_path / "use_capdec")
    lib_dir = os.path.dirname(_build.build())
    r = subprocess.run([gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(c_file), "-o", exe,
                        "-L", lib_dir, "-lcapdec", "-Wl,-rpath," + lib_dir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    n, rc, msg = out.stdout.strip().split(" ", 2)
    assert int(n) == len(syms)
    if not torch.cuda.is_available():
        assert int(rc) != 0 and ("CUDA" in msg or "no CPU fallback" in msg)


def test_caption_lengths_is_the_position_after_the_last_token():
    """`lengths` of capdec_train_fwd (include/capdec.h): [start, w_1 .. w_L, stop, pad ..] -> L + 2
    (preprocess/ai_challenge_tokenizer.py:81-86); rows t >= lengths - 1 carry no target (model:163-171)."""
    import numpy as np
    cap = torch.tensor([[2, 7, 9, 3, 0, 0],      # L = 2
                        [2, 3, 0, 0, 0, 0],      # empty sentence
                        [2, 5, 6, 7, 8, 3],      # fills the row
                        [0, 0, 0, 0, 0, 0],      # fully padded sample
                        [2, 5, 0, 6, 3, 0]])     # a hole inside the caption stays inside
    ln = M.caption_lengths(cap, 0)
    assert ln.dtype == np.int32 and ln.tolist() == [4, 2, 6, 0, 5]
    assert M.caption_lengths(cap.numpy(), 0).tolist() == [4, 2, 6, 0, 5]
    # on captions without holes the live rows of the packed head are exactly the rows the loss mask keeps
    cfg = M.DecoderConfig()
    caps = M.synth.make_captions(cfg, 256, 20260)
    T = cfg.sentence_length - 1
    live = np.clip(M.caption_lengths(caps, cfg.padding_idx).astype(np.int64) - 1, 0, T).sum()
    assert live == int((caps[:, 1:] != cfg.padding_idx).sum())
    assert 0.5 < live / (256 * T) < 0.65          # bench.py: "57 % of the B x 34 rows carry a target"


def test_parameter_versions_follow_in_place_updates():
    """The engine re-binds its operand copies when the flat parameter buffer's version moved: the registry is
    keyed by the buffer, bumped by Adam / ParamStore.load / checkpoint loads."""
    from myimagecaptioningmodel_b200 import _capi
    a, b = torch.zeros(8), torch.zeros(8)
    va, vb = _capi.version_of(a), _capi.version_of(b)
    _capi.bump_version(a)
    assert _capi.version_of(a) == va + 1 and _capi.version_of(b) == vb
    _capi.bump_version(a)
    _capi.bump_version(b)
    assert _capi.version_of(a) == va + 2 and _capi.version_of(b) == vb + 1
    # the product's optimizer refuses host buffers (there is no CPU path to update)
    with pytest.raises(ValueError):
        M.Adam([(a, torch.zeros(8))], 1e-3)
