package com.roklenarcic.util.strings.gpu;

import java.io.IOException;

import com.roklenarcic.util.strings.MapMatchListener;
import com.roklenarcic.util.strings.ReadableMatchListener;
import com.roklenarcic.util.strings.StringMap;
import com.roklenarcic.util.strings.threshold.Thresholder;

/** Drop-in for com.roklenarcic.util.strings.WholeWordLongestMatchMap (WholeWordLongestMatchMap.java:21-52,54,183): the six constructor overloads. */
public class WholeWordLongestMatchMap<T> extends GpuMatcher<T> implements StringMap<T> {
    private final boolean[] wordChars;

    public WholeWordLongestMatchMap(final Iterable<String> keywords, final Iterable<? extends T> values, boolean caseSensitive) {
        this(keywords, values, caseSensitive, AcGpuNative.wordChars(0, null, null), 0);
    }

    public WholeWordLongestMatchMap(final Iterable<String> keywords, final Iterable<? extends T> values, boolean caseSensitive,
            char[] wordCharacters) {
        this(keywords, values, caseSensitive, AcGpuNative.wordChars(1, wordCharacters, null), 0);
    }

    public WholeWordLongestMatchMap(final Iterable<String> keywords, final Iterable<? extends T> values, boolean caseSensitive,
            char[] wordCharacters, boolean[] toggleFlags) {
        this(keywords, values, caseSensitive, AcGpuNative.wordChars(2, wordCharacters, toggleFlags), 0);
    }

    /** The Thresholder overloads: it only shapes the reference's node objects and is ignored. */
    public WholeWordLongestMatchMap(final Iterable<String> keywords, final Iterable<? extends T> values, boolean caseSensitive,
            final Thresholder thresholdStrategy) {
        this(keywords, values, caseSensitive);
    }

    public WholeWordLongestMatchMap(final Iterable<String> keywords, final Iterable<? extends T> values, boolean caseSensitive,
            char[] wordCharacters, final Thresholder thresholdStrategy) {
        this(keywords, values, caseSensitive, wordCharacters);
    }

    public WholeWordLongestMatchMap(final Iterable<String> keywords, final Iterable<? extends T> values, boolean caseSensitive,
            char[] wordCharacters, boolean[] toggleFlags, final Thresholder thresholdStrategy) {
        this(keywords, values, caseSensitive, wordCharacters, toggleFlags);
    }

    private WholeWordLongestMatchMap(final Iterable<String> keywords, final Iterable<? extends T> values, boolean caseSensitive,
            boolean[] wordChars, int unused) {
        super(AcGpuNative.WHOLEWORDLONGEST, keywords, values, caseSensitive, wordChars);
        this.wordChars = wordChars;
    }

    public void match(final Readable haystack, final ReadableMatchListener<T> listener) throws IOException {
        matchReadable(haystack, listener);
    }

    public void match(final String haystack, final MapMatchListener<T> listener) {
        matchMap(haystack, listener);
    }

    /** getWordChars() - WholeWordLongestMatchSet.java:180 / WholeWordLongestMatchMap.java:308 */
    public boolean[] getWordChars() {
        return wordChars;
    }
}
