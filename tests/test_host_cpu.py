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
