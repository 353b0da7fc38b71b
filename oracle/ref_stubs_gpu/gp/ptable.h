#pragma once
/* TEST INFRASTRUCTURE -- stand-in for gp/ptable.h (common.h:5), the table printer of the LSU "gp" library.  One line per row,
 * "name=value" pairs: enough to read the reference's own numbers (t/us, GFlops, Ord, tm ...) off its stdout. */
#include <cstdio>
#include <string>
struct pTable {
  FILE* f;
  int num_lines = 0;
  bool in_row = false;
  explicit pTable(FILE* fp = stdout) : f(fp ? fp : stdout) {}
  template <class... A> void entry(const char* name, const char* fmt, A... a) {
    std::fprintf(f, " %s=", name);
    std::fprintf(f, fmt, a...);
  }
  void entry(const char* name, const char* fmt, const std::string& s) { std::fprintf(f, " %s=", name); std::fprintf(f, fmt, s.c_str()); }
  void header_span_start(const char*) {}
  void header_span_end() {}
  template <class... A> void header_span(const char*, A...) {}
  void row_start() { in_row = true; std::fprintf(f, "ROW"); }
  void row_end() { if (in_row) { std::fprintf(f, "\n"); ++num_lines; in_row = false; } }
};
struct pTable_Row {
  pTable& t;
  explicit pTable_Row(pTable& tt) : t(tt) { t.row_start(); }
  ~pTable_Row() { t.row_end(); }
};
