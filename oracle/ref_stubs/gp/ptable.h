#pragma once
/* stand-in for gp/ptable.h (common.h:5): the statistics table printer, a no-op here */
#include <cstdio>
struct pTable {
  explicit pTable(FILE* = nullptr) {}
  template <class... A> void entry(const char*, const char*, A...) {}
};
struct pTable_Row { explicit pTable_Row(pTable&) {} };
