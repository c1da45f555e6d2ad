/*
 * atmrt_fmt.h -- Rust's `{}` (Display) formatting of an f64, for the three text dumpers of the reference
 * (ray_path.rs:97-103, elev_profile.rs:62-64, atm_printer.rs:37-46 print with `{}`): the shortest digit string that
 * round-trips, laid out WITHOUT an exponent whatever the magnitude, no trailing ".0" (10.0 prints as "10"),
 * "NaN", "inf", "-inf", "-0". Header-only C++17 (std::to_chars yields the shortest digits).
 */
#ifndef ATMRT_FMT_H
#define ATMRT_FMT_H
#ifdef __cplusplus
#include <charconv>
#include <cmath>
#include <cstdlib>
#include <string>

inline std::string atmrt_fmt_f64(double v) {
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v < 0 ? "-inf" : "inf";
    char buf[64];
    auto res = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::scientific);  // shortest round-trip: d[.ddd]e[+-]xx
    std::string s(buf, res.ptr);
    std::string out;
    size_t i = 0;
    if (s[i] == '-') out += '-', ++i;
    const size_t e = s.find('e', i);
    std::string digits;
    for (size_t j = i; j < e; ++j)
        if (s[j] != '.') digits += s[j];
    const int exp10 = atoi(s.c_str() + e + 1);  // value = d.ddd * 10^exp10
    while (digits.size() > 1 && digits.back() == '0') digits.pop_back();
    const int point = exp10 + 1;  // digits before the decimal point
    const int n = (int)digits.size();
    if (digits == "0") return out + "0";
    if (point <= 0) {
        out += "0.";
        out.append((size_t)(-point), '0');
        out += digits;
    } else if (point >= n) {
        out += digits;
        out.append((size_t)(point - n), '0');
    } else {
        out += digits.substr(0, (size_t)point) + "." + digits.substr((size_t)point);
    }
    return out;
}
#endif
#endif /* ATMRT_FMT_H */
