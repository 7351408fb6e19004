/*
 * acgpu.hpp — C++17 host-side mirror of the reference's public API above the C ABI (include/acgpu.h).
 *
 * The reference (RokLenarcic/AhoCorasick) is compiled Java; this image has no JDK, so the host layer a user of the
 * reference would program against is restated here in C++ with the SAME class names, constructor argument order,
 * listener contract (true continues, false stops) and error behaviour.  Citations are relative to
 * src/main/java/com/roklenarcic/util/strings/ of the reference.  Header-only; link with -lacgpu.
 *
 *   Java                                              here
 *   ------------------------------------------------  ---------------------------------------------------------
 *   String haystack                                   acgpu::String (std::u16string — UTF-16 code units, Java char)
 *   Iterable<String> keywords                         any range of String / const char16_t* / std::optional<String>
 *                                                     (nullptr / nullopt = Java null: skipped, consumes a value)
 *   Iterable<? extends T> values                      any range of T; the matcher keeps copies (AhoCorasickMap.java:50)
 *   SetMatchListener / MapMatchListener<T> /          abstract classes of the same names, or any callable with the
 *   ReadableMatchListener<T>                          same arguments returning bool
 *   java.lang.Readable                                acgpu::Readable { int read(char16_t* dst, int capacity) } (-1 = EOF)
 *   IllegalArgumentException                          acgpu::IllegalArgumentException (std::invalid_argument)
 *   NullPointerException for null haystack/listener   not expressible (references)
 *   Thresholder / RangeNodeThreshold                  accepted and ignored (results-neutral)
 *
 * No CPU fallback: every constructor and match() needs a CUDA device and throws acgpu::Error(ACGPU_ENODEVICE)
 * without one.  Thread model as in the reference: a constructed matcher is immutable, concurrent match() is safe,
 * listener calls happen synchronously on the calling thread, in the reference's order.
 */
#ifndef ACGPU_HPP
#define ACGPU_HPP

#include <cmath>
#include <cstdint>
#include <functional>
#include <iterator>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <unordered_set>
#include <utility>
#include <vector>

#include "acgpu.h"

namespace acgpu {

using String = std::u16string;

/** java.lang.IllegalArgumentException as thrown by the WholeWord constructors (WholeWordMatchSet.java:149-153). */
class IllegalArgumentException : public std::invalid_argument {
   public:
    explicit IllegalArgumentException(const std::string &m) : std::invalid_argument(m) {}
};

/** Any other failure of the C ABI (code = ACGPU_E*). */
class Error : public std::runtime_error {
   public:
    Error(int code_, const std::string &m) : std::runtime_error(m), code(code_) {}
    const int code;
};

/** java.io.IOException for Readable implementations that want to fail (StringMap.java:6 `throws IOException`);
 *  it propagates through match(Readable, listener) untouched. */
class IOException : public std::runtime_error {
   public:
    explicit IOException(const std::string &m) : std::runtime_error(m) {}
};

inline void check(int rc) {
    if (rc == ACGPU_OK) return;
    const char *m = acgpu_last_error();
    std::string msg = m ? m : "";
    if (rc == ACGPU_EILLEGALARG) throw IllegalArgumentException(msg);
    throw Error(rc, "acgpu error " + std::to_string(rc) + ": " + msg);
}

// ------------------------------------------------------------------------------------------------ listeners

/** SetMatchListener.java:3-8. */
struct SetMatchListener {
    virtual ~SetMatchListener() = default;
    virtual bool match(const String &haystack, int startPosition, int endPosition) = 0;
};

/** MapMatchListener.java:3-8. */
template <class T>
struct MapMatchListener {
    virtual ~MapMatchListener() = default;
    virtual bool match(const String &haystack, int startPosition, int endPosition, const T &value) = 0;
};

/** ReadableMatchListener.java:4-8 — values only, no positions. */
template <class T>
struct ReadableMatchListener {
    virtual ~ReadableMatchListener() = default;
    virtual bool match(const T &value) = 0;
};

/** java.lang.Readable: fill dst with up to `capacity` chars, return how many, -1 at end of input. */
struct Readable {
    virtual ~Readable() = default;
    virtual int read(char16_t *dst, int capacity) = 0;
};

/** java.io.StringReader. */
class StringReader : public Readable {
   public:
    explicit StringReader(String s) : s_(std::move(s)) {}
    int read(char16_t *dst, int capacity) override {
        if (at_ >= s_.size()) return -1;
        size_t n = std::min(s_.size() - at_, (size_t)capacity);
        s_.copy(dst, n, at_);
        at_ += n;
        return (int)n;
    }
    size_t position() const { return at_; }

   private:
    String s_;
    size_t at_ = 0;
};

// ------------------------------------------------------------------------------------------------ thresholders

/** threshold/Thresholder.java:3-5 — shapes the reference's node objects only, never results; ignored here. */
struct Thresholder {
    virtual ~Thresholder() = default;
    virtual bool isOverThreshold(int nodeSize, int nodeLevel, int keyIntervalSize) const = 0;
};

/** threshold/RangeNodeThreshold.java:7-29 (kept so that user code constructing one still compiles and behaves). */
class RangeNodeThreshold : public Thresholder {
   public:
    RangeNodeThreshold() : RangeNodeThreshold(1, 1, 0.65, 2) {}
    RangeNodeThreshold(double exponent, double linearFactor, double maxValue, double constantFactor)
        : exponent_(exponent), linearFactor_(linearFactor), maxValue_(maxValue), constantFactor_(constantFactor) {}
    bool isOverThreshold(int nodeSize, int nodeLevel, int keyIntervalSize) const override {
        if (keyIntervalSize <= 8) return true;
        int charArrayCost = (nodeSize >> 2) + 3;
        return nodeSize + charArrayCost >
               keyIntervalSize * (maxValue_ - linearFactor_ / std::pow(constantFactor_ + nodeLevel, exponent_));
    }

   private:
    double exponent_, linearFactor_, maxValue_, constantFactor_;
};

// ------------------------------------------------------------------------------------------------ word characters

/** WordCharacters.java:4-63. */
struct WordCharacters {
    using Flags = std::vector<uint8_t>;  // the reference's boolean[65536]

    /** generateWordCharsFlags() — WordCharacters.java:6-16: isLetterOrDigit plus '-' and '_'. */
    static Flags generateWordCharsFlags() { return make(0, nullptr, nullptr, 0); }
    /** generateWordCharsFlags(char[]) — WordCharacters.java:18-24: only the given chars. */
    static Flags generateWordCharsFlags(const std::vector<char16_t> &wordCharacters) {
        return make(1, wordCharacters.data(), nullptr, (int)wordCharacters.size());
    }
    /** generateWordCharsFlags(char[], boolean[]) — WordCharacters.java:26-39: default table with toggles. */
    static Flags generateWordCharsFlags(const std::vector<char16_t> &characters, const std::vector<bool> &toggleFlags) {
        if (toggleFlags.size() < characters.size())
            throw std::out_of_range("toggleFlags shorter than characters");  // ArrayIndexOutOfBoundsException in Java
        std::vector<uint8_t> t(characters.size());
        for (size_t i = 0; i < t.size(); i++) t[i] = toggleFlags[i] ? 1 : 0;
        return make(2, characters.data(), t.data(), (int)characters.size());
    }
    /** trim(String, boolean[]) — WordCharacters.java:41-62. */
    static String trim(const String &keyword, const Flags &wordChars) {
        size_t b = 0, e = keyword.size();
        while (b < e && !wordChars[(uint16_t)keyword[b]]) b++;
        while (e > b && !wordChars[(uint16_t)keyword[e - 1]]) e--;
        return b < e ? keyword.substr(b, e - b) : keyword;
    }

   private:
    static Flags make(int mode, const char16_t *chars, const uint8_t *toggles, int n) {
        Flags out(65536);
        static const uint16_t none16 = 0;
        static const uint8_t none8 = 0;
        check(acgpu_word_chars(mode, chars ? reinterpret_cast<const uint16_t *>(chars) : &none16,
                               toggles ? toggles : &none8, n, out.data()));
        return out;
    }
};

// ------------------------------------------------------------------------------------------------ plumbing

namespace detail {

/** Keywords flattened the way acgpu_create_from_keywords takes them. */
struct Packed {
    std::vector<uint16_t> chars;
    std::vector<int64_t> offsets{0};
    std::vector<uint8_t> is_null;
    void add(const String &k) {
        chars.insert(chars.end(), k.begin(), k.end());
        offsets.push_back((int64_t)chars.size());
        is_null.push_back(0);
    }
    void add(const char16_t *k) {
        if (k) return add(String(k));
        add_null();
    }
    void add(const std::optional<String> &k) {
        if (k) return add(*k);
        add_null();
    }
    void add(std::nullptr_t) { add_null(); }
    void add_null() {
        offsets.push_back((int64_t)chars.size());
        is_null.push_back(1);
    }
    int64_t size() const { return (int64_t)is_null.size(); }
};

/** RAII acgpu_result. */
struct Result {
    acgpu_result r{0, nullptr, nullptr};
    Result() = default;
    Result(const Result &) = delete;
    Result &operator=(const Result &) = delete;
    ~Result() { acgpu_free_result(&r); }
    int64_t size() const { return r.n; }
    int start(int64_t i) const { return r.pos[2 * i]; }
    int end(int64_t i) const { return r.pos[2 * i + 1]; }
    uint32_t value(int64_t i) const { return r.val[i]; }
};

/** RAII acgpu_matches: the match stream of a Set in either wire format (records, or per-char hit masks when dense). */
struct Matches {
    acgpu_matches r{};
    Matches() = default;
    Matches(const Matches &) = delete;
    Matches &operator=(const Matches &) = delete;
    ~Matches() { acgpu_free_matches(&r); }
    int64_t size() const { return r.n; }
    bool masks() const { return r.kind == ACGPU_MATCHES_MASKS; }
};

/** Owns the device automaton; shared by the Set and Map façades. */
class Handle {
   public:
    Handle(int family, const Packed &kw, int64_t n_values, bool caseSensitive, const WordCharacters::Flags *wordChars,
           int device)
        : family_(family) {
        static const uint16_t none16 = 0;
        static const uint8_t none8 = 0;
        check(acgpu_create_from_keywords(family, kw.chars.empty() ? &none16 : kw.chars.data(), kw.offsets.data(),
                                         kw.is_null.empty() ? &none8 : kw.is_null.data(), kw.size(), n_values,
                                         caseSensitive ? 1 : 0, wordChars ? wordChars->data() : nullptr, device, &h_));
    }
    Handle(const Handle &) = delete;
    Handle &operator=(const Handle &) = delete;
    ~Handle() {
        if (h_) acgpu_destroy(h_);
    }
    uint64_t raw() const { return h_; }
    int family() const { return family_; }
    int charBufferSize() const {
        int64_t nodes, bytes;
        int32_t classes, maxLen, cbs;
        check(acgpu_info(h_, &nodes, &classes, &maxLen, &cbs, &bytes));
        return cbs;
    }
    void match(const String &hay, Result &out) const {
        if (hay.size() > 0x7fffffffu) throw Error(ACGPU_EINVAL, "haystack longer than a Java String");
        check(acgpu_match_utf16(h_, reinterpret_cast<const uint16_t *>(hay.data()), (int32_t)hay.size(), &out.r));
    }

    void match(const String &hay, Matches &out) const {
        if (hay.size() > 0x7fffffffu) throw Error(ACGPU_EINVAL, "haystack longer than a Java String");
        check(acgpu_match_utf16_compact(h_, reinterpret_cast<const uint16_t *>(hay.data()), (int32_t)hay.size(), &out.r));
    }

    /**
     * Replay of a Set stream in either wire format; emit(start, end) returns the listener's answer.  Hit masks are
     * expanded lazily, in the reference's order (end ascending, longest first - AhoCorasickSet.java:522-535): bit t of
     * masks[q] = a keyword of length 16 - t ends with char q.  Only the AhoCorasick family produces masks, so the
     * Shortest quirks never meet them.
     */
    template <class Emit>
    void replay(const Matches &m, size_t n_chars, Emit emit) const {
        if (!m.masks()) {
            const int32_t *pos = m.r.pos;
            const int64_t n = m.r.n;
            if (family_ == ACGPU_SHORTEST) {
                for (int64_t i = 0; i < n; i++) {
                    if ((size_t)pos[2 * i + 1] == n_chars) {
                        emit(pos[2 * i], pos[2 * i + 1]);
                        return;
                    }
                    if (!emit(pos[2 * i], pos[2 * i + 1])) {
                        emit(pos[2 * i], pos[2 * i + 1]);
                        return;
                    }
                }
                return;
            }
            for (int64_t i = 0; i < n; i++)
                if (!emit(pos[2 * i], pos[2 * i + 1])) return;
            return;
        }
        const uint16_t *mk = m.r.masks;
        for (int64_t q = 0; q < m.r.n_chars; q++) {
            uint32_t w = mk[q];
            while (w) {
                const int t = __builtin_ctz(w);
                w &= w - 1u;
                if (!emit((int)(q + 1 - (16 - t)), (int)(q + 1))) return;
            }
        }
    }

    /**
     * Replay of one match(String, listener) call.  emit(i) delivers record i and returns the listener's answer.
     * Shortest (ShortestMatchSet.java:196-226 and the Map twin): a `false` breaks out of the scan loop and falls into
     * the post-loop emit, which delivers the same match once more (quirk Q1); the match ending at haystack.length() is
     * only emitted post-loop and its return value is ignored (Q2).  Every other family stops at the first false.
     */
    template <class Emit>
    void replay(const Result &rec, size_t n_chars, Emit emit) const {
        const int64_t n = rec.size();
        if (family_ == ACGPU_SHORTEST) {
            for (int64_t i = 0; i < n; i++) {
                if ((size_t)rec.end(i) == n_chars) {
                    emit(i);
                    return;
                }
                if (!emit(i)) {
                    emit(i);
                    return;
                }
            }
            return;
        }
        for (int64_t i = 0; i < n; i++)
            if (!emit(i)) return;
    }

    /**
     * match(Readable, ReadableMatchListener) — StringMap.java:6.  The reader is drained in charBufferSize fills exactly
     * like the reference (AhoCorasickMap.java:53,213-219); fills are batched into device blocks and pushed through
     * acgpu_stream_*; the ordered value indices of every block go to emit(valueIdx).  ShortestMatchMap re-delivers a
     * match that ends exactly on a fill boundary and is followed by more input (quirk Q4, ShortestMatchMap.java:241-249).
     * Observable difference from the reference: the reader is consumed up to one block ahead of the listener calls.
     * blockChars = 0: blocks start at 64 Ki chars (an early-stopping caller over-reads little) and double up to 16 Mi
     * chars (the fixed cost of a feed is amortised); > 0: fixed block size.
     */
    template <class Emit>
    void matchReadable(Readable &in, Emit emit, size_t blockChars = 0) const {
        const bool adaptive = blockChars == 0;
        const size_t maxBlock = size_t(1) << 24;
        if (adaptive) blockChars = size_t(1) << 16;
        const int cbs = charBufferSize();
        struct Stream {
            uint64_t s = 0;
            ~Stream() {
                if (s) acgpu_stream_end(s, nullptr);
            }
        } st;
        check(acgpu_stream_begin(h_, &st.s));
        const bool shortest = family_ == ACGPU_SHORTEST;
        // the listener sees values only: leave the positions on the device, except for the Q4 replay of ShortestMatchMap
        check(acgpu_stream_set_values_only(st.s, shortest ? 0 : 1));
        std::unordered_set<int64_t> boundaries;
        int64_t n_read = 0;
        std::vector<char16_t> block(blockChars + (size_t)cbs);
        auto deliver = [&](const Result &rec) -> bool {
            for (int64_t i = 0; i < rec.size(); i++) {
                if (!emit(rec.value(i))) return false;
                if (shortest) {
                    int64_t e = rec.end(i);
                    if (e < n_read && boundaries.count(e) && !emit(rec.value(i))) return false;
                }
            }
            return true;
        };
        bool eof = false;
        while (!eof) {
            size_t got = 0;
            if (block.size() < blockChars + (size_t)cbs) block.resize(blockChars + (size_t)cbs);
            while (got < blockChars) {
                int k = in.read(block.data() + got, cbs);
                if (k < 0) {
                    eof = true;
                    break;
                }
                got += (size_t)k;
                n_read += k;
                if (shortest && k > 0) boundaries.insert(n_read);
                if (k == 0) break;  // a Readable may legally return 0; hand over what we have
            }
            if (got) {
                Result rec;
                check(acgpu_stream_feed(st.s, reinterpret_cast<const uint16_t *>(block.data()), (int32_t)got, &rec.r));
                if (!deliver(rec)) return;
            }
            if (adaptive) blockChars = std::min(2 * blockChars, maxBlock);
        }
        Result rec;
        uint64_t s = st.s;
        st.s = 0;
        check(acgpu_stream_end(s, &rec.r));
        deliver(rec);
    }

   private:
    uint64_t h_ = 0;
    int family_;
};

template <class Range>
Packed pack(const Range &keywords) {
    Packed p;
    for (const auto &k : keywords) p.add(k);
    return p;
}

/** Maps zip keywords with values and stop at the shorter (AhoCorasickMap.java:32-34). */
template <class T, class KRange, class VRange>
Packed zip(const KRange &keywords, const VRange &values, std::vector<T> &kept) {
    Packed p;
    auto ki = std::begin(keywords);
    auto vi = std::begin(values);
    for (; ki != std::end(keywords) && vi != std::end(values); ++ki, ++vi) {
        p.add(*ki);
        kept.push_back(*vi);
    }
    return p;
}

}  // namespace detail

// ------------------------------------------------------------------------------------------------ StringSet / StringMap

/** StringSet.java:3-5. */
class StringSet {
   public:
    virtual ~StringSet() = default;
    StringSet(StringSet &&) = default;  // matchers own their device automaton: movable, not copyable
    StringSet &operator=(StringSet &&) = default;

    /** match(String haystack, SetMatchListener listener) — StringSet.java:4. */
    void match(const String &haystack, SetMatchListener &listener) const {
        match(haystack, [&](const String &h, int s, int e) { return listener.match(h, s, e); });
    }
    /** Same, with any callable bool(const String&, int start, int end). */
    template <class F, class = std::enable_if_t<std::is_invocable_r_v<bool, F, const String &, int, int>>>
    void match(const String &haystack, F &&listener) const {
        detail::Matches rec;
        h_->match(haystack, rec);
        h_->replay(rec, haystack.size(), [&](int s, int e) { return (bool)listener(haystack, s, e); });
    }

   protected:
    template <class Range>
    StringSet(int family, const Range &keywords, bool caseSensitive, const WordCharacters::Flags *wc, int device)
        : h_(std::make_unique<detail::Handle>(family, detail::pack(keywords), -1, caseSensitive, wc, device)) {}
    std::unique_ptr<detail::Handle> h_;
};

/** StringMap.java:5-9. */
template <class T>
class StringMap {
   public:
    virtual ~StringMap() = default;
    StringMap(StringMap &&) = default;  // movable, not copyable
    StringMap &operator=(StringMap &&) = default;

    /** match(String haystack, MapMatchListener<T> listener) — StringMap.java:8. */
    void match(const String &haystack, MapMatchListener<T> &listener) const {
        match(haystack, [&](const String &h, int s, int e, const T &v) { return listener.match(h, s, e, v); });
    }
    template <class F, class = std::enable_if_t<std::is_invocable_r_v<bool, F, const String &, int, int, const T &>>>
    void match(const String &haystack, F &&listener) const {
        detail::Result rec;
        h_->match(haystack, rec);
        h_->replay(rec, haystack.size(), [&](int64_t i) {
            return (bool)listener(haystack, rec.start(i), rec.end(i), values_[rec.value(i)]);
        });
    }

    /** match(Readable haystack, ReadableMatchListener<T> listener) — StringMap.java:6. */
    void match(Readable &haystack, ReadableMatchListener<T> &listener) const {
        match(haystack, [&](const T &v) { return listener.match(v); });
    }
    template <class F, class = std::enable_if_t<std::is_invocable_r_v<bool, F, const T &>>>
    void match(Readable &haystack, F &&listener) const {
        h_->matchReadable(haystack, [&](uint32_t vi) { return (bool)listener(values_[vi]); });
    }

   protected:
    template <class KRange, class VRange>
    StringMap(int family, const KRange &keywords, const VRange &values, bool caseSensitive,
              const WordCharacters::Flags *wc, int device) {
        detail::Packed p = detail::zip<T>(keywords, values, values_);
        h_ = std::make_unique<detail::Handle>(family, p, (int64_t)values_.size(), caseSensitive, wc, device);
    }
    std::vector<T> values_;  // index = the valueIdx the kernels report (last duplicate wins; first for Shortest)
    std::unique_ptr<detail::Handle> h_;
};

// ------------------------------------------------------------------------------------------------ the ten public classes

#define ACGPU_PLAIN_SET(NAME, FAMILY, DOC)                                                                         \
    DOC class NAME : public StringSet {                                                                             \
       public:                                                                                                      \
        template <class Range>                                                                                      \
        NAME(const Range &keywords, bool caseSensitive, int device = 0)                                             \
            : StringSet(FAMILY, keywords, caseSensitive, nullptr, device) {}                                        \
        template <class Range>                                                                                      \
        NAME(const Range &keywords, bool caseSensitive, const Thresholder &, int device = 0)                        \
            : StringSet(FAMILY, keywords, caseSensitive, nullptr, device) {}                                        \
        NAME(std::initializer_list<String> keywords, bool caseSensitive, int device = 0)                            \
            : StringSet(FAMILY, keywords, caseSensitive, nullptr, device) {}                                        \
    };

#define ACGPU_PLAIN_MAP(NAME, FAMILY, DOC)                                                                         \
    DOC template <class T>                                                                                          \
    class NAME : public StringMap<T> {                                                                              \
       public:                                                                                                      \
        template <class KRange, class VRange>                                                                       \
        NAME(const KRange &keywords, const VRange &values, bool caseSensitive, int device = 0)                      \
            : StringMap<T>(FAMILY, keywords, values, caseSensitive, nullptr, device) {}                             \
        template <class KRange, class VRange>                                                                       \
        NAME(const KRange &keywords, const VRange &values, bool caseSensitive, const Thresholder &, int device = 0) \
            : StringMap<T>(FAMILY, keywords, values, caseSensitive, nullptr, device) {}                             \
    };

/** AhoCorasickSet.java:11-252 — every (overlapping) occurrence, ordered by end, longest first. */
ACGPU_PLAIN_SET(AhoCorasickSet, ACGPU_AHOCORASICK, )
/** AhoCorasickMap.java:14-336. */
ACGPU_PLAIN_MAP(AhoCorasickMap, ACGPU_AHOCORASICK, )
/** LongestMatchSet.java:10-265 — leftmost-longest, non-overlapping. */
ACGPU_PLAIN_SET(LongestMatchSet, ACGPU_LONGEST, )
/** LongestMatchMap.java:14-361. */
ACGPU_PLAIN_MAP(LongestMatchMap, ACGPU_LONGEST, )
/** ShortestMatchSet.java:10-260 — earliest end, non-overlapping. */
ACGPU_PLAIN_SET(ShortestMatchSet, ACGPU_SHORTEST, )
/** ShortestMatchMap.java:14-373. */
ACGPU_PLAIN_MAP(ShortestMatchMap, ACGPU_SHORTEST, )

#undef ACGPU_PLAIN_SET
#undef ACGPU_PLAIN_MAP

/** The six constructor overloads of the WholeWord classes: (keywords, caseSensitive), (…, char[] wordCharacters),
 *  (…, char[] characters, boolean[] toggleFlags), each also with a trailing Thresholder
 *  (WholeWordMatchSet.java:16,21,27,33,38,43). */
#define ACGPU_WW_SET(NAME, FAMILY)                                                                                   \
    class NAME : public StringSet {                                                                                   \
       public:                                                                                                        \
        template <class Range>                                                                                        \
        NAME(const Range &keywords, bool caseSensitive, int device = 0)                                               \
            : NAME(keywords, caseSensitive, WordCharacters::generateWordCharsFlags(), device, 0) {}                   \
        template <class Range>                                                                                        \
        NAME(const Range &keywords, bool caseSensitive, const Thresholder &, int device = 0)                          \
            : NAME(keywords, caseSensitive, WordCharacters::generateWordCharsFlags(), device, 0) {}                   \
        template <class Range>                                                                                        \
        NAME(const Range &keywords, bool caseSensitive, const std::vector<char16_t> &wordCharacters, int device = 0)  \
            : NAME(keywords, caseSensitive, WordCharacters::generateWordCharsFlags(wordCharacters), device, 0) {}     \
        template <class Range>                                                                                        \
        NAME(const Range &keywords, bool caseSensitive, const std::vector<char16_t> &wordCharacters,                  \
             const Thresholder &, int device = 0)                                                                     \
            : NAME(keywords, caseSensitive, WordCharacters::generateWordCharsFlags(wordCharacters), device, 0) {}     \
        template <class Range>                                                                                        \
        NAME(const Range &keywords, bool caseSensitive, const std::vector<char16_t> &characters,                      \
             const std::vector<bool> &toggleFlags, int device = 0)                                                    \
            : NAME(keywords, caseSensitive, WordCharacters::generateWordCharsFlags(characters, toggleFlags), device,  \
                   0) {}                                                                                              \
        template <class Range>                                                                                        \
        NAME(const Range &keywords, bool caseSensitive, const std::vector<char16_t> &characters,                      \
             const std::vector<bool> &toggleFlags, const Thresholder &, int device = 0)                               \
            : NAME(keywords, caseSensitive, WordCharacters::generateWordCharsFlags(characters, toggleFlags), device,  \
                   0) {}                                                                                              \
        /** getWordChars() — WholeWordMatchSet.java:134. */                                                           \
        const WordCharacters::Flags &getWordChars() const { return wordChars_; }                                      \
                                                                                                                      \
       private:                                                                                                       \
        template <class Range>                                                                                        \
        NAME(const Range &keywords, bool caseSensitive, WordCharacters::Flags wc, int device, int)                    \
            : StringSet(FAMILY, keywords, caseSensitive, &wc, device), wordChars_(std::move(wc)) {}                   \
        WordCharacters::Flags wordChars_;                                                                             \
    };

#define ACGPU_WW_MAP(NAME, FAMILY)                                                                                   \
    template <class T>                                                                                                \
    class NAME : public StringMap<T> {                                                                                \
       public:                                                                                                        \
        template <class KRange, class VRange>                                                                         \
        NAME(const KRange &keywords, const VRange &values, bool caseSensitive, int device = 0)                        \
            : NAME(keywords, values, caseSensitive, WordCharacters::generateWordCharsFlags(), device, 0) {}           \
        template <class KRange, class VRange>                                                                         \
        NAME(const KRange &keywords, const VRange &values, bool caseSensitive, const Thresholder &, int device = 0)   \
            : NAME(keywords, values, caseSensitive, WordCharacters::generateWordCharsFlags(), device, 0) {}           \
        template <class KRange, class VRange>                                                                         \
        NAME(const KRange &keywords, const VRange &values, bool caseSensitive,                                        \
             const std::vector<char16_t> &wordCharacters, int device = 0)                                             \
            : NAME(keywords, values, caseSensitive, WordCharacters::generateWordCharsFlags(wordCharacters), device,   \
                   0) {}                                                                                              \
        template <class KRange, class VRange>                                                                         \
        NAME(const KRange &keywords, const VRange &values, bool caseSensitive,                                        \
             const std::vector<char16_t> &wordCharacters, const Thresholder &, int device = 0)                        \
            : NAME(keywords, values, caseSensitive, WordCharacters::generateWordCharsFlags(wordCharacters), device,   \
                   0) {}                                                                                              \
        template <class KRange, class VRange>                                                                         \
        NAME(const KRange &keywords, const VRange &values, bool caseSensitive,                                        \
             const std::vector<char16_t> &characters, const std::vector<bool> &toggleFlags, int device = 0)           \
            : NAME(keywords, values, caseSensitive, WordCharacters::generateWordCharsFlags(characters, toggleFlags),  \
                   device, 0) {}                                                                                      \
        template <class KRange, class VRange>                                                                         \
        NAME(const KRange &keywords, const VRange &values, bool caseSensitive,                                        \
             const std::vector<char16_t> &characters, const std::vector<bool> &toggleFlags, const Thresholder &,      \
             int device = 0)                                                                                          \
            : NAME(keywords, values, caseSensitive, WordCharacters::generateWordCharsFlags(characters, toggleFlags),  \
                   device, 0) {}                                                                                      \
        /** getWordChars() — WholeWordMatchMap.java:242. */                                                           \
        const WordCharacters::Flags &getWordChars() const { return wordChars_; }                                      \
                                                                                                                      \
       private:                                                                                                       \
        template <class KRange, class VRange>                                                                         \
        NAME(const KRange &keywords, const VRange &values, bool caseSensitive, WordCharacters::Flags wc, int device,  \
             int)                                                                                                     \
            : StringMap<T>(FAMILY, keywords, values, caseSensitive, &wc, device), wordChars_(std::move(wc)) {}        \
        WordCharacters::Flags wordChars_;                                                                             \
    };

/** WholeWordMatchSet.java:8-205 — a keyword equal to a maximal run of word characters; keywords are trimmed and must
 *  not hold non-word characters inside (IllegalArgumentException, :147-153). */
ACGPU_WW_SET(WholeWordMatchSet, ACGPU_WHOLEWORD)
/** WholeWordMatchMap.java:13-339. */
ACGPU_WW_MAP(WholeWordMatchMap, ACGPU_WHOLEWORD)
/** WholeWordLongestMatchSet.java:9-260 — longest keyword from every walk start that is followed by a non-word char;
 *  keywords may hold non-word chars ("as if"). */
ACGPU_WW_SET(WholeWordLongestMatchSet, ACGPU_WHOLEWORDLONGEST)
/** WholeWordLongestMatchMap.java:13-420. */
ACGPU_WW_MAP(WholeWordLongestMatchMap, ACGPU_WHOLEWORDLONGEST)

#undef ACGPU_WW_SET
#undef ACGPU_WW_MAP

}  // namespace acgpu
#endif
