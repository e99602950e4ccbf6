// TEST INFRASTRUCTURE ONLY (oracle/): stand-in for boost::algorithm::trim (src/options.cpp:241)
#pragma once
#include <string>
namespace boost {
namespace algorithm {
inline void trim(std::string& s) {
    const char* ws = " \t\r\n\f\v";
    size_t b = s.find_first_not_of(ws);
    if (b == std::string::npos) { s.clear(); return; }
    size_t e = s.find_last_not_of(ws);
    s = s.substr(b, e - b + 1);
}
}  // namespace algorithm
}  // namespace boost
