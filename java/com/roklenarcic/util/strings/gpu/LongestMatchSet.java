package com.roklenarcic.util.strings.gpu;

import com.roklenarcic.util.strings.SetMatchListener;
import com.roklenarcic.util.strings.StringSet;
import com.roklenarcic.util.strings.threshold.Thresholder;

/** Drop-in for com.roklenarcic.util.strings.LongestMatchSet (LongestMatchSet.java:15,19,192). */
public class LongestMatchSet extends GpuMatcher<Void> implements StringSet {
    public LongestMatchSet(final Iterable<String> keywords, boolean caseSensitive) {
        super(AcGpuNative.LONGEST, keywords, null, caseSensitive, null);
    }

    /** The Thresholder only shapes the reference's node objects; it never changes results and is ignored. */
    public LongestMatchSet(final Iterable<String> keywords, boolean caseSensitive, final Thresholder thresholdStrategy) {
        this(keywords, caseSensitive);
    }

    public void match(final String haystack, final SetMatchListener listener) {
        matchSet(haystack, listener);
    }
}
