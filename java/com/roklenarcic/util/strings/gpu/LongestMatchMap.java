package com.roklenarcic.util.strings.gpu;

import java.io.IOException;

import com.roklenarcic.util.strings.MapMatchListener;
import com.roklenarcic.util.strings.ReadableMatchListener;
import com.roklenarcic.util.strings.StringMap;
import com.roklenarcic.util.strings.threshold.Thresholder;

/** Drop-in for com.roklenarcic.util.strings.LongestMatchMap (LongestMatchMap.java:19,23,203,288). */
public class LongestMatchMap<T> extends GpuMatcher<T> implements StringMap<T> {
    public LongestMatchMap(final Iterable<String> keywords, final Iterable<? extends T> values, boolean caseSensitive) {
        super(AcGpuNative.LONGEST, keywords, values, caseSensitive, null);
    }

    public LongestMatchMap(final Iterable<String> keywords, final Iterable<? extends T> values, boolean caseSensitive,
            final Thresholder thresholdStrategy) {
        this(keywords, values, caseSensitive);
    }

    public void match(final Readable haystack, final ReadableMatchListener<T> listener) throws IOException {
        matchReadable(haystack, listener);
    }

    public void match(final String haystack, final MapMatchListener<T> listener) {
        matchMap(haystack, listener);
    }
}
