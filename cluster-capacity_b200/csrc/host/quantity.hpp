// quantity.hpp — resource.Quantity as the scheduler uses it: exact value (held in nano units), Value()/MilliValue()
// rounding UP (vendor/k8s.io/apimachinery/pkg/api/resource/quantity.go:812-834) and the canonical String() form
// (quantity.go CanonicalizeBytes) needed by the report's podRequirements.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>

namespace cch {

typedef __int128 i128;

struct Quantity {
  enum Format { DecimalSI, BinarySI, DecimalExponent };
  i128 nanos = 0;      // value * 1e9, exact (anything finer than nano is rounded up at parse time, like the reference)
  Format format = DecimalSI;

  static i128 pow10(int e) { i128 r = 1; while (e-- > 0) r *= 10; return r; }
  static i128 ceil_div(i128 a, i128 b) {  // b > 0
    i128 q = a / b, r = a % b;
    if (r != 0 && ((r > 0) == (b > 0))) q += 1;
    return q;
  }
  int64_t value() const { return (int64_t)ceil_div(nanos, pow10(9)); }        // Quantity.Value(): ceil
  int64_t milli_value() const { return (int64_t)ceil_div(nanos, pow10(6)); }  // Quantity.MilliValue(): ceil
  bool is_zero() const { return nanos == 0; }
  void add(const Quantity &o) { nanos += o.nanos; }   // q.Add(y): keeps q's format

  static Quantity parse(const std::string &s) {
    Quantity q;
    size_t p = 0, n = s.size();
    if (n == 0) throw std::runtime_error("quantity: empty");
    bool neg = false;
    if (s[p] == '+' || s[p] == '-') { neg = s[p] == '-'; p++; }
    i128 mant = 0;
    int frac_digits = 0;
    bool any = false, in_frac = false;
    for (; p < n; p++) {
      char c = s[p];
      if (c >= '0' && c <= '9') { mant = mant * 10 + (c - '0'); if (in_frac) frac_digits++; any = true; }
      else if (c == '.' && !in_frac) in_frac = true;
      else break;
    }
    if (!any) throw std::runtime_error("quantity: no digits in '" + s + "'");
    std::string suf = s.substr(p);
    int e10 = 0, e2 = 0;
    q.format = DecimalSI;
    if (!suf.empty() && (suf[0] == 'e' || suf[0] == 'E') && suf.size() > 1 &&
        (isdigit((unsigned char)suf[1]) || ((suf[1] == '+' || suf[1] == '-') && suf.size() > 2 && isdigit((unsigned char)suf[2])))) {
      e10 = atoi(suf.c_str() + 1);
      q.format = DecimalExponent;
    } else if (suf == "") e10 = 0;
    else if (suf == "n") e10 = -9; else if (suf == "u") e10 = -6; else if (suf == "m") e10 = -3;
    else if (suf == "k") e10 = 3; else if (suf == "M") e10 = 6; else if (suf == "G") e10 = 9;
    else if (suf == "T") e10 = 12; else if (suf == "P") e10 = 15; else if (suf == "E") e10 = 18;
    else if (suf == "Ki") { e2 = 10; q.format = BinarySI; } else if (suf == "Mi") { e2 = 20; q.format = BinarySI; }
    else if (suf == "Gi") { e2 = 30; q.format = BinarySI; } else if (suf == "Ti") { e2 = 40; q.format = BinarySI; }
    else if (suf == "Pi") { e2 = 50; q.format = BinarySI; } else if (suf == "Ei") { e2 = 60; q.format = BinarySI; }
    else throw std::runtime_error("quantity: bad suffix in '" + s + "'");
    // nanos = mant * 10^(e10 - frac_digits + 9) * 2^e2
    int e = e10 - frac_digits + 9;
    i128 v = mant;
    for (int i = 0; i < e2; i++) v *= 2;
    if (e >= 0) v *= pow10(e);
    else v = ceil_div(v, pow10(-e));   // finer than nano: round up
    q.nanos = neg ? -v : v;
    return q;
  }

  static std::string i128_str(i128 v) {
    if (v == 0) return "0";
    bool neg = v < 0;
    if (neg) v = -v;
    std::string s;
    while (v > 0) { s.insert(s.begin(), (char)('0' + (int)(v % 10))); v /= 10; }
    return neg ? "-" + s : s;
  }

  // Quantity.String(): canonical form in the quantity's own format (quantity.go CanonicalizeBytes)
  std::string str() const {
    if (nanos == 0) return "0";
    Format f = format;
    const i128 unit = pow10(9);
    if (f == BinarySI) {
      i128 a = nanos < 0 ? -nanos : nanos;
      if (a < 1024 * unit || nanos % unit != 0) f = DecimalSI;   // |q| < 1024 or fractional: decimal form
    }
    if (f == BinarySI) {
      i128 v = nanos / unit;
      static const char *suf[] = {"", "Ki", "Mi", "Gi", "Ti", "Pi", "Ei"};
      int k = 0;
      while (k < 6 && v % 1024 == 0) { v /= 1024; k++; }
      return i128_str(v) + suf[k];
    }
    // decimal: mantissa * 10^exp with exp a multiple of 3 and no trailing zeros absorbed beyond that
    i128 m = nanos;
    int exp = -9;
    while (m % 10 == 0) { m /= 10; exp++; }
    while (exp % 3 != 0) { m *= 10; exp--; }   // exp may be negative: C++ % keeps the sign, loop still terminates at a multiple of 3
    if (exp > 18) { m *= pow10(exp - 18); exp = 18; }
    if (f == DecimalExponent) return exp == 0 ? i128_str(m) : i128_str(m) + "e" + std::to_string(exp);
    const char *suf = "";
    switch (exp) {
      case -9: suf = "n"; break; case -6: suf = "u"; break; case -3: suf = "m"; break; case 0: suf = ""; break;
      case 3: suf = "k"; break; case 6: suf = "M"; break; case 9: suf = "G"; break; case 12: suf = "T"; break;
      case 15: suf = "P"; break; case 18: suf = "E"; break;
    }
    return i128_str(m) + suf;
  }
};

}  // namespace cch
