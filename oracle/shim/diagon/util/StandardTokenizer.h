// ORACLE-ONLY build shim (test infrastructure, never linked into the product).
//
// Shadows /root/reference/src/core/include/diagon/util/StandardTokenizer.h:6-8, whose only
// non-ASCII path needs ICU headers that this image does not have. Synthetic corpora are
// pre-tokenised ASCII ("t0000123 t0004567 ..."), so only the ASCII behaviour matters:
// split on characters that are not [A-Za-z0-9] (an apostrophe continues a token), lowercase.
#pragma once

#include <string>
#include <vector>

namespace diagon {
namespace util {

class StandardTokenizer {
public:
    static std::vector<std::string> tokenize(const std::string& text) {
        std::vector<std::string> out;
        std::string cur;
        auto alnum = [](unsigned char c) {
            return (c >= '0' && c <= '9') || (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z');
        };
        for (unsigned char c : text) {
            if (alnum(c) || (c == '\'' && !cur.empty())) {
                cur.push_back(static_cast<char>((c >= 'A' && c <= 'Z') ? c + ('a' - 'A') : c));
            } else if (!cur.empty()) {
                out.push_back(cur);
                cur.clear();
            }
        }
        if (!cur.empty()) out.push_back(cur);
        return out;
    }
};

}  // namespace util
}  // namespace diagon
