"""Deterministic synthetic workloads for the BASELINE.json configs.

Everything is a pure function of (seed, index) through SplitMix64, so numpy (tests, CPU oracle),
torch-on-GPU (bench: large corpora generated on the device, no multi-GB H2D) and any future JVM run
produce identical dictionaries and haystacks.  Not part of the product path.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np

MASK64 = (1 << 64) - 1
GOLDEN = 0x9E3779B97F4A7C15


def _mix_np(z: np.ndarray) -> np.ndarray:
    z = z.astype(np.uint64, copy=True)
    z ^= z >> np.uint64(30)
    z *= np.uint64(0xBF58476D1CE4E5B9)
    z ^= z >> np.uint64(27)
    z *= np.uint64(0x94D049BB133111EB)
    z ^= z >> np.uint64(31)
    return z


def hash_np(seed: int, idx: np.ndarray) -> np.ndarray:
    """SplitMix64 output number idx (0-based) of the stream seeded with `seed`."""
    with np.errstate(over="ignore"):
        state = (np.uint64(seed & MASK64) + (idx.astype(np.uint64) + np.uint64(1)) * np.uint64(GOLDEN))
        return _mix_np(state)


# ----------------------------------------------------------------------------- dictionaries

def make_keywords(n: int, seed: int, alphabet: str = "abcdefghijklmnopqrstuvwxyz", min_len: int = 3,
                  max_len: int = 12, nested_fraction: float = 0.0) -> List[str]:
    """n distinct keywords, chars uniform over `alphabet`, length uniform in [min_len, max_len] (re-drawn on
    collision, so short lengths saturate for huge n).  nested_fraction > 0 makes that share of the keywords
    prefixes / suffixes / infixes of other keywords (configs[2])."""
    alpha = np.array([ord(c) for c in alphabet], dtype=np.uint16)
    out: List[str] = []
    seen = set()
    counter = 0
    span = max_len - min_len + 1
    n_base = n if nested_fraction <= 0 else max(1, int(n * (1.0 - nested_fraction)))
    while len(out) < n_base:
        need = n_base - len(out)
        batch = int(need * 1.3) + 16
        idx = np.arange(counter, counter + batch * (max_len + 1), dtype=np.uint64)
        h = hash_np(seed, idx).reshape(batch, max_len + 1)
        counter += batch * (max_len + 1)
        lens = (h[:, 0] % np.uint64(span)).astype(np.int64) + min_len
        letters = alpha[(h[:, 1:] % np.uint64(alpha.size)).astype(np.int64)]
        raw = letters.astype("<u2").tobytes().decode("utf-16-le")
        for i in range(batch):
            w = raw[i * max_len:i * max_len + int(lens[i])]
            if w not in seen:
                seen.add(w)
                out.append(w)
                if len(out) == n_base:
                    break
    # nested keywords: deterministic substrings of earlier keywords
    j = 0
    while len(out) < n:
        hj = hash_np(seed ^ 0x5EED, np.arange(3 * j, 3 * j + 3, dtype=np.uint64))
        j += 1
        base = out[int(hj[0] % np.uint64(n_base))]
        if len(base) <= min_len:
            continue
        ln = min_len + int(hj[1] % np.uint64(len(base) - min_len))
        st = int(hj[2] % np.uint64(len(base) - ln + 1))
        w = base[st:st + ln]
        if w not in seen:
            seen.add(w)
            out.append(w)
    return out


def keywords_to_arrays(keywords: List[str]) -> Tuple[np.ndarray, np.ndarray]:
    """(chars uint16, offsets int64[n+1])"""
    lens = np.fromiter((len(k) for k in keywords), dtype=np.int64, count=len(keywords))
    offsets = np.zeros(len(keywords) + 1, np.int64)
    np.cumsum(lens, out=offsets[1:])
    chars = np.frombuffer("".join(keywords).encode("utf-16-le", "surrogatepass"), dtype=np.uint16)
    return chars, offsets


_ONSETS = ["", "b", "c", "d", "f", "g", "h", "j", "k", "l", "m", "n", "p", "r", "s", "t", "v", "w", "y", "z", "bl", "br", "ch", "cl", "cr",
           "dr", "fl", "fr", "gl", "gr", "pl", "pr", "qu", "sc", "sh", "sk", "sl", "sm", "sn", "sp", "st", "str", "sw", "th", "tr", "tw", "wh"]
_VOWELS = ["a", "e", "i", "o", "u", "a", "e", "i", "o", "ai", "ea", "ee", "ie", "io", "oa", "oo", "ou", "ue", "y"]
_CODAS = ["", "", "", "b", "ck", "d", "g", "l", "ll", "m", "n", "nd", "ng", "nt", "p", "r", "rd", "rs", "s", "ss", "st", "t", "th", "x"]
_SUFFIXES = ["", "", "", "", "s", "ed", "ing", "ly", "er", "est", "ness", "ment", "tion", "able", "'s", "n't", "'ll", "'re"]


def make_english_like(n: int = 236_000, seed: int = 1006, max_len: int = 24) -> List[str]:
    """n distinct English-LIKE words (README.md:130 of the reference quotes a 235 886-word English dictionary): 1-6
    syllables of onset + vowel + coda, common suffixes, apostrophes ("'s", "n't"), 1..max_len chars; 20 % Capitalised, 5 %
    UPPER CASE, the rest lower case - a case-sensitive dictionary over 53 symbols.  Deterministic in (n, seed)."""
    out: List[str] = ["a", "I", "an", "the", "of", "to", "in", "it", "is", "The", "A"]
    seen = set(out)
    counter = 0
    while len(out) < n:
        batch = 4096
        h = hash_np(seed, np.arange(counter, counter + batch * 24, dtype=np.uint64)).reshape(batch, 24)
        counter += batch * 24
        for row in h:
            k = 1 + int(row[0] % np.uint64(100)) * 6 // 100 if int(row[1] & np.uint64(3)) else 1 + int(row[0] % np.uint64(3))
            parts = []
            for j in range(k):
                parts.append(_ONSETS[int(row[2 + 3 * j] % np.uint64(len(_ONSETS)))])
                parts.append(_VOWELS[int(row[3 + 3 * j] % np.uint64(len(_VOWELS)))])
                parts.append(_CODAS[int(row[4 + 3 * j] % np.uint64(len(_CODAS)))])
            w = "".join(parts) + _SUFFIXES[int(row[20] % np.uint64(len(_SUFFIXES)))]
            if not w or len(w) > max_len or w[0] == "'":
                continue
            style = int(row[21] % np.uint64(20))
            if style < 4:
                w = w[0].upper() + w[1:]
            elif style == 4:
                w = w.upper()
            if w not in seen:
                seen.add(w)
                out.append(w)
                if len(out) == n:
                    break
    return out


ENGLISH_BLOCK = 1 << 20  # chars per independently generated text block
_SEPS = [" ", " ", " ", " ", " ", " ", ", ", ". ", "; ", "\n", " - ", "? "]


def _english_block(spec: "HaystackSpec", block: int) -> np.ndarray:
    """Block `block` of the infinite English-like text of spec: dictionary words (85 %) and out-of-dictionary words
    (15 %) separated by spaces and punctuation, exactly ENGLISH_BLOCK chars."""
    words = spec.words
    nw = len(words)
    per = ENGLISH_BLOCK // 3 + 64  # more tokens than can fit
    h = hash_np(spec.seed ^ 0xE791, np.arange(block * per * 2, (block + 1) * per * 2, dtype=np.uint64)).reshape(per, 2)
    pick = (h[:, 0] % np.uint64(nw)).astype(np.int64)
    # skew towards the head of the dictionary (short, frequent words)
    pick = np.where((h[:, 1] & np.uint64(3)) == 0, pick % 512, pick)
    oov = ((h[:, 1] >> np.uint64(8)) % np.uint64(100)) < 15
    sep = ((h[:, 1] >> np.uint64(16)) % np.uint64(len(_SEPS))).astype(np.int64)
    parts, total = [], 0
    for i in range(per):
        w = words[pick[i]]
        if oov[i]:
            w = w[::-1] + "q"
        parts.append(w)
        parts.append(_SEPS[sep[i]])
        total += len(w) + len(_SEPS[sep[i]])
        if total >= ENGLISH_BLOCK:
            break
    text = "".join(parts)
    assert len(text) >= ENGLISH_BLOCK
    return np.frombuffer(text[:ENGLISH_BLOCK].encode("utf-16-le"), dtype=np.uint16)


# ----------------------------------------------------------------------------- haystacks

BLOCK = 64  # one planted keyword per 64-char block
PUNCT = " ,.;:_()[]\"\n"
EXOTIC = "ÀÉÎÕÜàéîõüßΑΒΓΔΩαβγδωАБВГДабвгд"


class HaystackSpec:
    """style:
      'lower'  : a-z with spaces (p = 1/8)                                              (configs 0, 2, 4)
      'mixed'  : random case A-Za-z, 5 % Latin-1/Greek/Cyrillic letters, planted keywords in random case (config 1)
      'words'  : words separated by punctuation runs; 10 % of the words are keywords, 10 % keywords with an
                 extra word char glued on                                                (config 3)
      'english': English-like running text built from the dictionary's own words         (config 5, the "real dictionary")
    """

    def __init__(self, style: str, seed: int, keywords: List[str], plant: bool = True):
        self.style = style
        self.seed = seed
        self.plant = plant
        self.words = keywords if style == "english" else None
        self.kw_chars, self.kw_offsets = keywords_to_arrays(keywords)
        self.kw_lens = np.diff(self.kw_offsets)
        self.max_len = int(self.kw_lens.max()) if len(keywords) else 0


def _base_chars_np(spec: HaystackSpec, start: int, n: int) -> np.ndarray:
    idx = np.arange(start, start + n, dtype=np.uint64)
    h = hash_np(spec.seed, idx)
    letter = ((h >> np.uint64(8)) % np.uint64(26)).astype(np.uint16)
    if spec.style == "lower":
        out = letter + np.uint16(ord("a"))
        out[(h & np.uint64(7)) == 0] = ord(" ")
        return out
    if spec.style == "mixed":
        upper = ((h >> np.uint64(16)) & np.uint64(1)).astype(np.uint16)
        out = letter + np.uint16(ord("a")) - upper * np.uint16(32)
        out[(h & np.uint64(7)) == 0] = ord(" ")
        ex = np.array([ord(c) for c in EXOTIC], dtype=np.uint16)
        sel = ((h >> np.uint64(20)) % np.uint64(20)) == 0
        out[sel] = ex[((h[sel] >> np.uint64(32)) % np.uint64(ex.size)).astype(np.int64)]
        return out
    if spec.style == "words":
        alpha = np.array([ord(c) for c in "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789=-"], dtype=np.uint16)
        out = alpha[((h >> np.uint64(8)) % np.uint64(alpha.size)).astype(np.int64)]
        pun = np.array([ord(c) for c in PUNCT], dtype=np.uint16)
        sel = (h & np.uint64(7)) < 2
        out[sel] = pun[((h[sel] >> np.uint64(32)) % np.uint64(pun.size)).astype(np.int64)]
        return out
    raise ValueError(spec.style)


def make_haystack(spec: HaystackSpec, n: int, start: int = 0) -> np.ndarray:
    """chars [start, start+n) of the infinite haystack defined by spec (start must be a multiple of BLOCK)."""
    assert start % BLOCK == 0
    if spec.style == "english":
        b0, b1 = start // ENGLISH_BLOCK, (start + n + ENGLISH_BLOCK - 1) // ENGLISH_BLOCK
        text = np.concatenate([_english_block(spec, b) for b in range(b0, max(b1, b0 + 1))])
        return text[start - b0 * ENGLISH_BLOCK:start - b0 * ENGLISH_BLOCK + n].copy()
    out = _base_chars_np(spec, start, n)
    if not spec.plant or spec.kw_lens.size == 0 or spec.max_len > BLOCK - 4:
        return out
    nb = (n + BLOCK - 1) // BLOCK
    b = np.arange(start // BLOCK, start // BLOCK + nb, dtype=np.uint64)
    hb = hash_np(spec.seed ^ 0xB10C, b)
    k = ((hb >> np.uint64(1)) % np.uint64(spec.kw_lens.size)).astype(np.int64)
    ln = spec.kw_lens[k]
    glue = np.zeros(nb, dtype=bool)
    if spec.style == "words":
        # the planted word gets a separator on both sides; some get one extra word char glued on (must NOT match)
        mode = ((hb >> np.uint64(40)) % np.uint64(10)).astype(np.int64)
        planted = mode < 2
        glue = mode == 1
    else:
        planted = np.ones(nb, dtype=bool)
    room = BLOCK - 3 - ln
    off = 1 + ((hb >> np.uint64(20)) % np.uint64(BLOCK)).astype(np.int64) % np.maximum(room, 1)
    pos0 = (b.astype(np.int64) - start // BLOCK) * BLOCK + off
    case_bits = (hb >> np.uint64(44)).astype(np.uint64)
    for j in range(spec.max_len):
        m = planted & (j < ln)
        p = pos0[m] + j
        ok = p < n
        ch = spec.kw_chars[spec.kw_offsets[k[m]] + j]
        if spec.style == "mixed":
            up = ((case_bits[m] >> np.uint64(j)) & np.uint64(1)).astype(bool) & (ch >= ord("a")) & (ch <= ord("z"))
            ch = np.where(up, ch - 32, ch).astype(np.uint16)
        out[p[ok]] = ch[ok]
    if spec.style == "words":
        before = pos0[planted] - 1
        after = pos0[planted] + ln[planted] + glue[planted].astype(np.int64)
        for p in (before, after):
            ok = (p >= 0) & (p < n)
            out[p[ok]] = ord(" ")
        p = pos0[glue] + ln[glue]
        ok = p < n
        out[p[ok]] = ord("x")
    return out


# ----------------------------------------------------------------------------- same generator in torch (device side)

def _i64(x: int) -> int:
    x &= MASK64
    return x - (1 << 64) if x >= (1 << 63) else x


def _lsr(z, k: int):
    return (z >> k) & ((1 << (64 - k)) - 1)


def hash_torch(seed: int, idx):
    """SplitMix64 in wrapping int64 arithmetic; bit-identical to hash_np."""
    z = idx * _i64(GOLDEN) + _i64((seed & MASK64) + GOLDEN)
    z = z ^ _lsr(z, 30)
    z = z * _i64(0xBF58476D1CE4E5B9)
    z = z ^ _lsr(z, 27)
    z = z * _i64(0x94D049BB133111EB)
    z = z ^ _lsr(z, 31)
    return z


def make_haystack_torch(spec: HaystackSpec, n: int, start: int = 0, device="cuda", out=None, chunk: int = 1 << 25):
    """make_haystack() evaluated with torch ops on `device` ('lower' and 'mixed' styles), chunk by chunk so a
    10^9-char haystack needs no large temporaries.  Returns an int16 tensor holding the uint16 code units."""
    import torch
    assert start % BLOCK == 0 and chunk % BLOCK == 0
    if spec.style == "english":
        # text blocks are built on the host (word-level generator); long haystacks tile a 32-block period
        period = 32 * ENGLISH_BLOCK
        if out is None:
            out = torch.empty(n, dtype=torch.int16, device=device)
        if n <= period:
            out.copy_(torch.from_numpy(make_haystack(spec, n, start).view(np.int16)).to(device))
            return out
        assert start % period == 0
        base = torch.from_numpy(make_haystack(spec, period, 0).view(np.int16)).to(device)
        for c0 in range(0, n, period):
            m = min(period, n - c0)
            out[c0:c0 + m] = base[:m]
        return out
    assert spec.style in ("lower", "mixed", "words")
    if out is None:
        out = torch.empty(n, dtype=torch.int16, device=device)
    plant = spec.plant and spec.kw_lens.size > 0 and spec.max_len <= BLOCK - 4
    if plant:
        kw_chars = torch.from_numpy(spec.kw_chars.astype(np.int64)).to(device)
        kw_offsets = torch.from_numpy(spec.kw_offsets.astype(np.int64)).to(device)
        kw_lens = torch.from_numpy(spec.kw_lens.astype(np.int64)).to(device)
    ex = torch.tensor([ord(c) for c in EXOTIC], dtype=torch.int64, device=device)
    alpha_w = torch.tensor([ord(c) for c in "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789=-"],
                           dtype=torch.int64, device=device)
    pun = torch.tensor([ord(c) for c in PUNCT], dtype=torch.int64, device=device)
    for c0 in range(0, n, chunk):
        m = min(chunk, n - c0)
        idx = torch.arange(start + c0, start + c0 + m, dtype=torch.int64, device=device)
        h = hash_torch(spec.seed, idx)
        letter = _lsr(h, 8) % 26
        if spec.style == "lower":
            ch = letter + ord("a")
            ch = torch.where((h & 7) == 0, torch.full_like(ch, ord(" ")), ch)
        elif spec.style == "words":
            ch = alpha_w[_lsr(h, 8) % alpha_w.numel()]
            ch = torch.where((h & 7) < 2, pun[_lsr(h, 32) % pun.numel()], ch)
        else:
            upper = _lsr(h, 16) & 1
            ch = letter + ord("a") - upper * 32
            ch = torch.where((h & 7) == 0, torch.full_like(ch, ord(" ")), ch)
            sel = (_lsr(h, 20) % 20) == 0
            ch = torch.where(sel, ex[_lsr(h, 32) % ex.numel()], ch)
        del h, letter, idx
        if plant:
            nb = (m + BLOCK - 1) // BLOCK
            b = torch.arange((start + c0) // BLOCK, (start + c0) // BLOCK + nb, dtype=torch.int64, device=device)
            hb = hash_torch(spec.seed ^ 0xB10C, b)
            k = _lsr(hb, 1) % int(spec.kw_lens.size)
            ln = kw_lens[k]
            room = BLOCK - 3 - ln
            off = 1 + (_lsr(hb, 20) % BLOCK) % torch.clamp(room, min=1)
            pos0 = (b - (start + c0) // BLOCK) * BLOCK + off
            case_bits = _lsr(hb, 44)
            if spec.style == "words":
                mode = _lsr(hb, 40) % 10
                planted = mode < 2
                glue = mode == 1
            else:
                planted = torch.ones_like(ln, dtype=torch.bool)
                glue = torch.zeros_like(planted)
            for j in range(spec.max_len):
                msk = planted & (j < ln)
                p = pos0[msk] + j
                kc = kw_chars[kw_offsets[k[msk]] + j]
                if spec.style == "mixed":
                    up = ((case_bits[msk] >> j) & 1).bool() & (kc >= ord("a")) & (kc <= ord("z"))
                    kc = torch.where(up, kc - 32, kc)
                ok = p < m
                ch[p[ok]] = kc[ok]
            if spec.style == "words":
                before = pos0[planted] - 1
                after = pos0[planted] + ln[planted] + glue[planted].to(torch.int64)
                for p in (before, after):
                    ok = (p >= 0) & (p < m)
                    ch[p[ok]] = ord(" ")
                p = pos0[glue] + ln[glue]
                ok = p < m
                ch[p[ok]] = ord("x")
        out[c0:c0 + m] = ch.to(torch.int16)  # wraps values >= 0x8000 into the same 16 bits
        del ch
    return out


# ----------------------------------------------------------------------------- the five configs

def config(idx: int, scale: float = 1.0):
    """(family, is_map, case_sensitive, keywords, HaystackSpec, n_chars, extra) of BASELINE.json configs[idx];
    `scale` shrinks the dictionary and haystack for tests."""
    if idx == 0:
        kws = make_keywords(max(10, int(1000 * scale)), 1001)
        return dict(family="ahocorasick", is_map=False, cs=True, keywords=kws,
                    spec=HaystackSpec("lower", 2001, kws), n=int(8_000_000 * scale))
    if idx == 1:
        kws = make_keywords(max(10, int(100_000 * scale)), 1002)
        return dict(family="ahocorasick", is_map=True, cs=False, keywords=kws,
                    spec=HaystackSpec("mixed", 2002, kws), n=int(500_000_000 * scale))
    if idx == 2:
        kws = make_keywords(max(10, int(100_000 * scale)), 1003, nested_fraction=0.25)
        return dict(family="longest", is_map=True, cs=True, keywords=kws,
                    spec=HaystackSpec("lower", 2003, kws), n=int(2_000_000_000 * scale))
    if idx == 3:
        kws = make_keywords(max(10, int(50_000 * scale)), 1004,
                            alphabet="abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ0123456789=-")
        return dict(family="wholeword", is_map=False, cs=True, keywords=kws,
                    spec=HaystackSpec("words", 2004, kws), n=int(2_000_000_000 * scale),
                    word_chars=(["_", "="], [False, True]))
    if idx == 4:
        kws = make_keywords(max(10, int(1_000_000 * scale)), 1005)
        return dict(family="ahocorasick", is_map=False, cs=True, keywords=kws,
                    spec=HaystackSpec("lower", 2005, kws), n=int(1_000_000_000 * scale))
    if idx == 5:
        # not a BASELINE config: the "real dictionary" workload of VERDICT r01 (README.md:130 of the reference: 235 886 English
        # words) - 236 000 English-like words, 1-24 chars, mixed case and apostrophes (53 symbols), case-sensitive
        kws = make_english_like(max(50, int(236_000 * scale)), 1006)
        return dict(family="ahocorasick", is_map=False, cs=True, keywords=kws,
                    spec=HaystackSpec("english", 2006, kws), n=int(1_000_000_000 * scale))
    raise ValueError(idx)
