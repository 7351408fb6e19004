"""CPU-only checks of the drop-in boundary and the host logic: the C-ABI library loads and exports every symbol
include/acgpu.h declares, the product never touches the oracle, dictionary validation happens before any device work,
and without a CUDA device every matcher fails loudly (there is no CPU matching path)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "acgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(acgpu_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_header_symbol():
    from ahocorasick_b200 import _lib
    syms = _header_symbols()
    assert len(syms) >= 14
    assert sorted(_lib.EXPORTS) == syms
    lib = C.CDLL(os.path.join(ROOT, "ahocorasick_b200", "libacgpu.so"))
    for s in syms:
        assert getattr(lib, s) is not None, s


def test_product_does_not_import_oracle():
    """oracle/ is test infrastructure: nothing under ahocorasick_b200/ (Python or C++) may name it."""
    pkg = os.path.join(ROOT, "ahocorasick_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in os.path.relpath(dirpath, pkg).split(os.sep)[0:1]:
            continue
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                continue
            text = open(os.path.join(dirpath, f), errors="ignore").read()
            assert not re.search(r"\boracle\b|liboracle|ac_oracle", text), os.path.join(dirpath, f)


def test_word_chars_tables_match_the_oracle():
    """acgpu_word_chars (WordCharacters.java:6-39) against the oracle's restatement, all three constructors."""
    from ahocorasick_b200 import WordCharacters
    from oracle import oracle as ora
    assert np.array_equal(WordCharacters.generateWordCharsFlags().astype(np.uint8), ora.word_chars(0))
    assert np.array_equal(WordCharacters.generateWordCharsFlags(["_", "="]).astype(np.uint8), ora.word_chars(1, ["_", "="]))
    assert np.array_equal(WordCharacters.generateWordCharsFlags(["_", "="], [False, True]).astype(np.uint8),
                          ora.word_chars(2, ["_", "="], [False, True]))
    wc = WordCharacters.generateWordCharsFlags()
    assert WordCharacters.trim("  ,as if. ", wc) == "as if" and WordCharacters.trim(",;", wc) == ",;"


def test_dictionary_validation_precedes_device_work():
    """IllegalArgumentException for WholeWord keywords with inner non-word chars (WholeWordMatchSet.java:149-153) is
    raised by the host-side flattening, with the reference's message, whether or not a GPU exists."""
    import ahocorasick_b200 as ac
    with pytest.raises(ac.IllegalArgumentException, match="as if contains non-word characters."):
        ac.WholeWordMatchSet(["as", " as if "], True)
    with pytest.raises(ac.IllegalArgumentException):
        ac.WholeWordMatchMap(["a,b"], [1], False)
    with pytest.raises(ac.IllegalArgumentException):
        ac.WholeWordMatchSet(["abc"], True, ["_", "="])  # custom-only word chars: alphanumerics are no word chars


def test_no_device_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    import ahocorasick_b200 as ac
    from ahocorasick_b200 import _lib
    for cls, args in ((ac.AhoCorasickSet, (["ab"], True)), (ac.LongestMatchMap, (["ab"], [1], True)),
                      (ac.WholeWordLongestMatchSet, (["as if"], False))):
        with pytest.raises(ac.AcgpuError) as e:
            cls(*args)
        assert e.value.code == _lib.ENODEVICE


def test_keyword_packing_fast_path_equals_per_keyword_path():
    """_pack_keywords joins all-str dictionaries and encodes once; the layout must equal the per-keyword path (which a
    None keyword forces), including empty keywords, non-BMP characters (UTF-16 length != len(str)) and lone surrogates."""
    from ahocorasick_b200.matchers import _pack_keywords
    rng = np.random.default_rng(7)
    cases = [["he", "she", "hers"], ["", "", ""], ["a\U0001F600b", "", "xy"], ["é", "日本語", "\ud800x"], [],
             ["".join(chr(int(c)) for c in rng.integers(1, 0xD7FF, size=int(rng.integers(0, 9)))) for _ in range(500)]]
    for kws in cases:
        chars, offsets, is_null, n = _pack_keywords(kws)
        chars2, offsets2, is_null2, n2 = _pack_keywords(list(kws) + [None])
        assert n == len(kws) and n2 == n + 1 and is_null2[-1] == 1 and not is_null[:n].any()
        assert np.array_equal(offsets, offsets2[:-1])
        assert np.array_equal(chars[:offsets[-1]], chars2[:offsets[-1]])
        for i, k in enumerate(kws):
            assert chars[offsets[i]:offsets[i + 1]].tobytes() == k.encode("utf-16-le", "surrogatepass")
    it = _pack_keywords(k for k in ["ab", "c"])  # any Iterable, like the Java constructors
    assert it[3] == 2 and it[1].tolist() == [0, 2, 3]
