package com.roklenarcic.util.strings.gpu;

import java.io.IOException;

import com.roklenarcic.util.strings.MapMatchListener;
import com.roklenarcic.util.strings.ReadableMatchListener;
import com.roklenarcic.util.strings.StringMap;
import com.roklenarcic.util.strings.threshold.Thresholder;

/** Drop-in for com.roklenarcic.util.strings.WholeWordMatchMap (WholeWordMatchMap.java:21-53,55,155): the six constructor overloads. */
public class WholeWordMatchMap<T> extends GpuMatcher<T> implements StringMap<T> {
    private final boolean[] wordChars;

    public WholeWordMatchMap(final Iterable<String> keywords, final Iterable<? extends T> values, boolean caseSensitive) {
        this(keywords, values, caseSensitive, AcGpuNative.wordChars(0, null, null), 0);
    }

    public WholeWordMatchMap(final Iterable<String> keywords, final Iterable<? extends T> values, boolean caseSensitive,
            char[] wordCharacters) {
        this(keywords, values, caseSensitive, AcGpuNative.wordChars(1, wordCharacters, null), 0);
    }

    public WholeWordMatchMap(final Iterable<String> keywords, final Iterable<? extends T> values, boolean caseSensitive,
            char[] wordCharacters, boolean[] toggleFlags) {
        this(keywords, values, caseSensitive, AcGpuNative.wordChars(2, wordCharacters, toggleFlags), 0);
    }

    /** The Thresholder overloads: it only shapes the reference's node objects and is ignored. */
    public WholeWordMatchMap(final Iterable<String> keywords, final Iterable<? extends T> values, boolean caseSensitive,
            final Thresholder thresholdStrategy) {
        this(keywords, values, caseSensitive);
    }

    public WholeWordMatchMap(final Iterable<String> keywords, final Iterable<? extends T> values, boolean caseSensitive,
            char[] wordCharacters, final Thresholder thresholdStrategy) {
        this(keywords, values, caseSensitive, wordCharacters);
    }

    public WholeWordMatchMap(final Iterable<String> keywords, final Iterable<? extends T> values, boolean caseSensitive,
            char[] wordCharacters, boolean[] toggleFlags, final Thresholder thresholdStrategy) {
        this(keywords, values, caseSensitive, wordCharacters, toggleFlags);
    }

    private WholeWordMatchMap(final Iterable<String> keywords, final Iterable<? extends T> values, boolean caseSensitive,
            boolean[] wordChars, int unused) {
        super(AcGpuNative.WHOLEWORD, keywords, values, caseSensitive, wordChars);
        this.wordChars = wordChars;
    }

    public void match(final Readable haystack, final ReadableMatchListener<T> listener) throws IOException {
        matchReadable(haystack, listener);
    }

    public void match(final String haystack, final MapMatchListener<T> listener) {
        matchMap(haystack, listener);
    }

    /** getWordChars() - WholeWordMatchSet.java:134 / WholeWordMatchMap.java:242 */
    public boolean[] getWordChars() {
        return wordChars;
    }
}
