package com.roklenarcic.util.strings.gpu;

import java.io.IOException;
import java.nio.CharBuffer;
import java.util.ArrayList;
import java.util.Iterator;
import java.util.List;

import com.roklenarcic.util.strings.MapMatchListener;
import com.roklenarcic.util.strings.ReadableMatchListener;
import com.roklenarcic.util.strings.SetMatchListener;

/**
 * Shared body of the eight drop-in classes: flatten the Iterables into the arrays acgpu_create_from_keywords takes,
 * call the native scan, replay the ordered records to the listener with the reference's early-stop behaviour
 * (ahocorasick_b200/matchers.py is the tested twin of this class).
 */
abstract class GpuMatcher<T> implements AutoCloseable {
    private long handle; // 0 after close()
    private final int family;
    private final Object[] values; // null for Sets; index = the valueIdx the kernels report

    GpuMatcher(int family, Iterable<String> keywords, Iterable<? extends T> valuesIn, boolean caseSensitive,
            boolean[] wordChars) {
        this.family = family;
        List<String> kws = new ArrayList<String>();
        List<Object> vals = valuesIn == null ? null : new ArrayList<Object>();
        Iterator<String> ki = keywords.iterator();
        Iterator<? extends T> vi = valuesIn == null ? null : valuesIn.iterator();
        // Maps zip keywords with values and stop at the shorter (AhoCorasickMap.java:32); null/empty keywords are
        // skipped by the builder but still consume a value (AhoCorasickMap.java:33-36).
        while (ki.hasNext() && (vi == null || vi.hasNext())) {
            kws.add(ki.next());
            if (vi != null) {
                vals.add(vi.next());
            }
        }
        int total = 0;
        for (String k : kws) {
            total += k == null ? 0 : k.length();
        }
        char[] chars = new char[Math.max(total, 1)];
        long[] offsets = new long[kws.size() + 1];
        byte[] isNull = new byte[Math.max(kws.size(), 1)];
        int at = 0;
        for (int i = 0; i < kws.size(); i++) {
            String k = kws.get(i);
            offsets[i] = at;
            if (k == null) {
                isNull[i] = 1;
            } else {
                k.getChars(0, k.length(), chars, at);
                at += k.length();
            }
        }
        offsets[kws.size()] = at;
        this.values = vals == null ? null : vals.toArray();
        this.handle = AcGpuNative.create(family, chars, offsets, isNull, kws.size(), vals == null ? -1 : vals.size(),
                caseSensitive, wordChars, 0);
    }

    /** Releases the device tables; idempotent.  Matching after close() throws IllegalStateException. */
    public synchronized void close() {
        long h = handle;
        handle = 0;
        if (h != 0) {
            AcGpuNative.destroy(h);
        }
    }

    private long live() {
        long h = handle;
        if (h == 0) {
            throw new IllegalStateException("matcher is closed");
        }
        return h;
    }

    /** StringSet.match(String, SetMatchListener) — StringSet.java:4 */
    protected void matchSet(String haystack, SetMatchListener listener) {
        // acgpu_match_utf16_compact: {int[] pos, int[] valueIdx, char[] masks}; dense AhoCorasickSet streams arrive as
        // per-char hit masks (2 B/char) and are expanded lazily here, in the reference's order (end ascending, longest
        // first - AhoCorasickSet.java:522-535): bit t of masks[q] = a keyword of length 16 - t ends with char q
        Object[] r = AcGpuNative.matchCompact(live(), haystack);
        if (r[2] != null) {
            char[] masks = (char[]) r[2];
            for (int q = 0; q < masks.length; q++) {
                for (int w = masks[q]; w != 0; w &= w - 1) {
                    int t = Integer.numberOfTrailingZeros(w);
                    if (!listener.match(haystack, q + 1 - (16 - t), q + 1)) {
                        return;
                    }
                }
            }
            return;
        }
        int[] pos = (int[]) r[0];
        int n = pos.length / 2;
        for (int i = 0; i < n; i++) {
            boolean more = listener.match(haystack, pos[2 * i], pos[2 * i + 1]);
            if (family == AcGpuNative.SHORTEST) {
                // ShortestMatchSet.java:196-226: the match ending at haystack.length() is emitted post-loop (return
                // value ignored); a `false` on any other match falls into the post-loop emit and delivers it again.
                if (pos[2 * i + 1] == haystack.length()) {
                    return;
                }
                if (!more) {
                    listener.match(haystack, pos[2 * i], pos[2 * i + 1]);
                    return;
                }
            } else if (!more) {
                return;
            }
        }
    }

    /** StringMap.match(String, MapMatchListener) — StringMap.java:8 */
    @SuppressWarnings("unchecked")
    protected void matchMap(String haystack, MapMatchListener<T> listener) {
        Object[] r = AcGpuNative.match(live(), haystack);
        int[] pos = (int[]) r[0], val = (int[]) r[1];
        for (int i = 0; i < val.length; i++) {
            T v = (T) values[val[i]];
            boolean more = listener.match(haystack, pos[2 * i], pos[2 * i + 1], v);
            if (family == AcGpuNative.SHORTEST) {
                if (pos[2 * i + 1] == haystack.length()) {
                    return;
                }
                if (!more) {
                    listener.match(haystack, pos[2 * i], pos[2 * i + 1], v);
                    return;
                }
            } else if (!more) {
                return;
            }
        }
    }

    /**
     * StringMap.match(Readable, ReadableMatchListener) — StringMap.java:6.  Reads the Readable in charBufferSize fills
     * exactly like the reference (AhoCorasickMap.java:213-219), batches the fills into device blocks (64 Ki chars doubling to 16 Mi) and replays
     * the ordered value indices of every block.  ShortestMatchMap re-delivers a match that ends exactly on a fill boundary
     * and is followed by more input (quirk Q4, ShortestMatchMap.java:241-249); the replay knows every fill boundary.
     * Tested twins of this method: include/acgpu.hpp (detail::Handle::matchReadable) and ahocorasick_b200/streaming.py.
     */
    @SuppressWarnings("unchecked")
    protected void matchReadable(Readable haystack, ReadableMatchListener<T> listener) throws IOException {
        final int cbs = AcGpuNative.charBufferSize(live());
        int blockChars = 1 << 16;            // device blocks start small (early stop over-reads little) ...
        final int maxBlockChars = 1 << 24;   // ... and double up to 16 Mi chars (the fixed cost of a feed is amortised)
        final boolean shortest = family == AcGpuNative.SHORTEST;
        final java.util.HashSet<Long> boundaries = new java.util.HashSet<Long>();
        long s = AcGpuNative.streamBegin(live());
        AcGpuNative.streamValuesOnly(s, !shortest); // the listener sees values only; ShortestMatchMap's Q4 replay needs the ends
        boolean ended = false;
        try {
            char[] block = new char[blockChars + cbs];
            long nRead = 0;
            boolean eof = false;
            while (!eof) {
                int got = 0;
                if (block.length < blockChars + cbs) {
                    block = new char[blockChars + cbs];
                }
                while (got < blockChars) {
                    int k = haystack.read(CharBuffer.wrap(block, got, cbs));
                    if (k < 0) {
                        eof = true;
                        break;
                    }
                    got += k;
                    nRead += k;
                    if (shortest && k > 0) {
                        boundaries.add(nRead);
                    }
                    if (k == 0) {
                        break;
                    }
                }
                if (got > 0 && !replayValues(AcGpuNative.streamFeed(s, block, got), listener, shortest, boundaries, nRead)) {
                    return;
                }
                blockChars = Math.min(2 * blockChars, maxBlockChars);
            }
            ended = true;
            replayValues(AcGpuNative.streamEnd(s), listener, shortest, boundaries, nRead);
        } finally {
            if (!ended) {
                AcGpuNative.streamEnd(s);
            }
        }
    }

    @SuppressWarnings("unchecked")
    private boolean replayValues(Object[] r, ReadableMatchListener<T> listener, boolean shortest,
            java.util.HashSet<Long> boundaries, long nRead) {
        int[] pos = (int[]) r[0], val = (int[]) r[1];
        for (int i = 0; i < val.length; i++) {
            T v = (T) values[val[i]];
            if (!listener.match(v)) {
                return false;
            }
            if (shortest) {
                long e = pos[2 * i + 1];
                if (e < nRead && boundaries.contains(e) && !listener.match(v)) {
                    return false;
                }
            }
        }
        return true;
    }
}
