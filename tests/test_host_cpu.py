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


def _fingerprint(family, kws, n_values=-1, cs=True, wc=None):
    from ahocorasick_b200 import _lib
    from ahocorasick_b200.matchers import _pack_keywords
    chars, offsets, is_null, n = _pack_keywords(kws)
    fp = C.c_uint64(0)
    _lib.check(_lib.lib().acgpu_build_fingerprint(family, chars.ctypes.data, offsets.ctypes.data, is_null.ctypes.data, n, n_values,
                                                  1 if cs else 0, wc.ctypes.data if wc is not None else None, C.byref(fp), None))
    return fp.value


def test_sharded_builder_equals_serial_builder(monkeypatch):
    """Dictionary flattening: the concurrent per-first-class insert (csrc/trie_insert.hpp, used for
    50 000+ keywords) must give bit-identical tables to the serial insert - every family, Set and Map, both trie
    directions, duplicates (last wins / first wins for Shortest), None keywords, case folding, wide alphabets."""
    rng = np.random.default_rng(11)

    def words(n, alpha, lo, hi):
        return ["".join(rng.choice(list(alpha), size=int(rng.integers(lo, hi + 1)))) for _ in range(n)]

    dicts = [
        words(3000, "abcdefghijklmnopqrstuvwxyz", 3, 12),
        words(2000, "ab", 1, 16) + words(50, "ab", 1, 3) * 3,          # heavy sharing and duplicates
        words(1500, "ABCdef0123=-", 1, 9) + [None, "", "Zz"],           # case folding, nulls
        [chr(c) for c in range(0x20, 0x2000)] + words(200, "αβγЖж", 2, 5),  # wide alphabet
        [],
    ]
    seen = set()
    for kws in dicts:
        for family in range(5):
            for n_values in (-1, len(kws)):
                for cs in (True, False):
                    try:
                        monkeypatch.setenv("ACGPU_BUILDER", "serial")
                        a = _fingerprint(family, kws, n_values, cs)
                    except Exception as e:   # WholeWord refuses keywords with inner non-word chars: both must refuse
                        monkeypatch.setenv("ACGPU_BUILDER", "sharded")
                        with pytest.raises(type(e)):
                            _fingerprint(family, kws, n_values, cs)
                        continue
                    monkeypatch.setenv("ACGPU_BUILDER", "sharded")
                    b = _fingerprint(family, kws, n_values, cs)
                    assert a == b, (family, n_values, cs, len(kws))
                    seen.add(a)
    assert len(seen) > 40   # the fingerprint does depend on the dictionary, the family and the flags


def test_builder_default_mode_is_deterministic_at_scale(monkeypatch):
    """100 000 keywords take the sharded path by default; two builds and the forced serial build agree."""
    import workloads as W
    kws = W.config(1)["keywords"]
    monkeypatch.delenv("ACGPU_BUILDER", raising=False)
    a = _fingerprint(0, kws, len(kws), False)
    b = _fingerprint(0, kws, len(kws), False)
    monkeypatch.setenv("ACGPU_BUILDER", "serial")
    c = _fingerprint(0, kws, len(kws), False)
    assert a == b == c


def test_trie_insert_unit_program():
    """tests/cpp/trie_insert_test.cpp: 720 random cases, sharded (1, 4 and 13 threads) == serial on the raw arrays."""
    import subprocess
    out_dir = os.path.join(ROOT, "tests", "cpp", "build")
    os.makedirs(out_dir, exist_ok=True)
    exe = os.path.join(out_dir, "trie_insert_test")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-pthread", os.path.join(ROOT, "tests", "cpp", "trie_insert_test.cpp"),
                           "-o", exe])
    out = subprocess.run([exe, "4"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and " 0 failures" in out.stdout, out.stdout + out.stderr


def test_header_is_plain_c_and_links_from_c():
    """The drop-in boundary is a C ABI: include/acgpu.h compiles as strict C99 (no C++ or torch types in the signatures),
    every declared entry point links from a C program, and the host-only calls behave (tests/cpp/c_abi_check.c)."""
    import subprocess
    out_dir = os.path.join(ROOT, "tests", "cpp", "build")
    os.makedirs(out_dir, exist_ok=True)
    exe = os.path.join(out_dir, "c_abi_check")
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "c_abi_check.c"), "-L" + os.path.join(ROOT, "ahocorasick_b200"),
                           "-lacgpu", "-Wl,-rpath," + os.path.join(ROOT, "ahocorasick_b200"), "-o", exe])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "c abi ok: %d entry points" % len(_header_symbols()) in r.stdout, (r.returncode, r.stdout, r.stderr)


def test_builder_tables_match_committed_fingerprints(monkeypatch):
    """The CUDA kernels read the flattened tables, and only a GPU run can prove a new layout right.  This guard fails when
    a change alters the tables of seeded dictionaries (tests/golden/builder_fingerprints.json, generated by
    tests/golden/make_builder_fingerprints.py): re-run the GPU parity suite, then regenerate the file."""
    import json
    import sys
    monkeypatch.delenv("ACGPU_BUILDER", raising=False)
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_builder_fingerprints as g
    want = json.load(open(g.OUT))
    got = g.compute()
    assert sorted(got) == sorted(want)
    diff = [k for k in want if got[k] != want[k]]
    assert not diff, "flattened tables changed for %d dictionaries, e.g. %s" % (len(diff), diff[:3])


def test_integration_doc_names_every_entry_point():
    """INTEGRATION.md maps every C entry point to the reference interface it replaces; a symbol added to include/acgpu.h
    must be documented there (and in the Java native class when the facade needs it)."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = [s for s in _header_symbols() if s not in doc and s not in ("acgpu_last_error", "acgpu_version", "acgpu_free_result",
                                                                          "acgpu_destroy")]
    assert not missing, missing
    for s in ("acgpu_last_error", "acgpu_free_result"):
        assert s in doc


def test_trie_descriptor_flattens_like_the_keyword_constructor():
    """acgpu_create(const acgpu_automaton_desc*) (host half, acgpu_desc_fingerprint): the tables derived from a flattened
    goto trie are those of acgpu_create_from_keywords for the dictionary the trie spells - Maps keep their value indices,
    Sets re-insert the keywords in trie order (which reproduces the state numbering); failure links, when given, are
    checked; malformed descriptors are refused."""
    import workloads as W
    from ahocorasick_b200 import _lib, trie_desc
    from ahocorasick_b200.matchers import _pack_keywords

    def fp_kw(family, kws, n_values, cs):
        chars, offsets, is_null, n = _pack_keywords(kws)
        fp = C.c_uint64(0)
        _lib.check(_lib.lib().acgpu_build_fingerprint(family, chars.ctypes.data, offsets.ctypes.data, is_null.ctypes.data, n, n_values,
                                                      int(cs), None, C.byref(fp), None))
        return fp.value

    kws = W.make_keywords(4000, 21) + ["a", "ab", "abc"]
    for family in range(5):
        for cs in (True, False):
            t = trie_desc.flatten_trie(family, kws + [None], list(range(len(kws) + 1)), cs)
            assert t.fingerprint() == fp_kw(family, kws + [None], len(kws) + 1, cs), (family, cs)
            t = trie_desc.flatten_trie(family, kws, None, cs)
            # the dictionary in trie order: one keyword per terminal state, ascending
            words = []
            for s in np.flatnonzero(t.terminal):
                u = []
                while s > 0:
                    u.append(int(t.edge_char[s]))
                    s = int(t.parent[s])
                words.append("".join(map(chr, reversed(u))))
            assert sorted(words) == sorted(set(kws))
            assert t.fingerprint() == fp_kw(family, words, -1, cs), (family, cs)
    t = trie_desc.flatten_trie(0, ["abcd", "bcd", "cd"], None, True)
    assert t.fail.tolist() == [0, 0, 5, 6, 7, 0, 8, 9, 0, 0]
    t.fail[3] = 2
    with pytest.raises(_lib.AcgpuError):
        t.fingerprint()
    t = trie_desc.flatten_trie(0, ["ab", "b"], [0, 1], True)
    t.value[:] = 1  # two states carry one value index
    with pytest.raises(_lib.AcgpuError):
        t.fingerprint()
    t = trie_desc.flatten_trie(0, ["ab", "b"], None, True)
    t.parent[1] = 2  # parents must precede children
    with pytest.raises(_lib.AcgpuError):
        t.fingerprint()
