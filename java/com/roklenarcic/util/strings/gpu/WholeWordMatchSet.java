package com.roklenarcic.util.strings.gpu;

import com.roklenarcic.util.strings.SetMatchListener;
import com.roklenarcic.util.strings.StringSet;
import com.roklenarcic.util.strings.threshold.Thresholder;

/**
 * Drop-in for com.roklenarcic.util.strings.WholeWordMatchSet (WholeWordMatchSet.java:16-45,47): the six constructor overloads; the word
 * character table comes from acgpu_word_chars (the reference's WordCharacters is package-private).
 */
public class WholeWordMatchSet extends GpuMatcher<Void> implements StringSet {
    private final boolean[] wordChars;

    public WholeWordMatchSet(final Iterable<String> keywords, boolean caseSensitive) {
        this(keywords, caseSensitive, AcGpuNative.wordChars(0, null, null), 0);
    }

    public WholeWordMatchSet(final Iterable<String> keywords, boolean caseSensitive, char[] wordCharacters) {
        this(keywords, caseSensitive, AcGpuNative.wordChars(1, wordCharacters, null), 0);
    }

    public WholeWordMatchSet(final Iterable<String> keywords, boolean caseSensitive, char[] wordCharacters, boolean[] toggleFlags) {
        this(keywords, caseSensitive, AcGpuNative.wordChars(2, wordCharacters, toggleFlags), 0);
    }

    /** The Thresholder overloads: it only shapes the reference's node objects and is ignored. */
    public WholeWordMatchSet(final Iterable<String> keywords, boolean caseSensitive, final Thresholder thresholdStrategy) {
        this(keywords, caseSensitive);
    }

    public WholeWordMatchSet(final Iterable<String> keywords, boolean caseSensitive, char[] wordCharacters,
            final Thresholder thresholdStrategy) {
        this(keywords, caseSensitive, wordCharacters);
    }

    public WholeWordMatchSet(final Iterable<String> keywords, boolean caseSensitive, char[] wordCharacters, boolean[] toggleFlags,
            final Thresholder thresholdStrategy) {
        this(keywords, caseSensitive, wordCharacters, toggleFlags);
    }

    private WholeWordMatchSet(final Iterable<String> keywords, boolean caseSensitive, boolean[] wordChars, int unused) {
        super(AcGpuNative.WHOLEWORD, keywords, null, caseSensitive, wordChars);
        this.wordChars = wordChars;
    }

    public void match(final String haystack, final SetMatchListener listener) {
        matchSet(haystack, listener);
    }

    /** getWordChars() - WholeWordMatchSet.java:134 / WholeWordMatchMap.java:242 (package-private in the reference; public here because the facade lives in its own package) */
    public boolean[] getWordChars() {
        return wordChars;
    }
}
