// The C++ usage example of INTEGRATION.md section 5, as a program.
//   g++ -std=c++17 -Iinclude examples/cpp_usage.cpp -Lahocorasick_b200 -lacgpu -Wl,-rpath,$PWD/ahocorasick_b200 -o cpp_usage
// Needs a CUDA device at run time (there is no CPU path); tests/test_cpp_host.py compiles it and runs it against the
// oracle-mocked C ABI on the CPU.  Expected output: "1 4 2", "2 4 1", "2 6 3", "v 2", "v 1", "v 3", "ww 1".
#include "acgpu.hpp"
#include <cstdio>
using namespace acgpu;                                   // String = std::u16string (Java char = UTF-16 code unit)
int main() {
std::vector<String> kws = {u"he", u"she", u"hers"};
AhoCorasickMap<int> m(kws, std::vector<int>{1, 2, 3}, /*caseSensitive=*/false);   // (keywords, values, caseSensitive)
m.match(u"uSHErs", [](const String&, int start, int end, const int& v) { std::printf("%d %d %d\n", start, end, v); return true; });   // true = continue
StringReader in(u"ushers");
m.match(in, [](const int& v) { std::printf("v %d\n", v); return true; });          // Readable overload: values only (StringMap.java:6)

struct Count : SetMatchListener {                        // or implement the interface, like in Java
    int n = 0;
    bool match(const String&, int, int) override { return ++n < 10; }   // false stops the scan
} count;
WholeWordMatchSet ww(kws, true, std::vector<char16_t>{u'_', u'='}, std::vector<bool>{false, true});
ww.match(u"he_she=hers", count);
std::printf("ww %d\n", count.n);
return 0;
}
