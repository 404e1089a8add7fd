// Deterministic, stateless TPC-H-shaped data generator, compiled for host AND device.
//
// The reference benchmarks on dbgen/tpchgen output (benchmarks/README.md:30-60) which is not
// available offline; SURVEY.md §8(d) therefore prescribes a counter-based generator that follows
// the TPC-H 4.2.3 value distributions and the reference's table schemas
// (benchmarks/src/bin/tpch.rs:960-1049: Int64 keys, Decimal128(15,2) money, Date32 dates, Utf8
// strings).  Every value is a pure function of (table, column, row index, scale factor), so any
// row range can be generated on any GPU or on the CPU (oracle) with identical bytes.
//
// This is synthetic-input infrastructure: it is not part of the query path.
#pragma once
#include <cstdint>

#include "hash.hpp"

namespace b200 {
namespace tpch {

enum Table : int { LINEITEM = 0, ORDERS, CUSTOMER, SUPPLIER, PART, PARTSUPP, NATION, REGION, N_TABLES };

// Column type tags for the generator (logical types are fixed by the reference schema).
enum ColKind : int { K_I64 = 0, K_I32, K_DEC, K_DATE, K_STR };

struct ColDef {
  const char* name;
  ColKind kind;
};

// Column order == benchmarks/src/bin/tpch.rs:960-1049
#define B200_TPCH_MAX_COLS 16
static const int kNumCols[N_TABLES] = {16, 9, 8, 7, 9, 5, 4, 3};
static const ColDef kCols[N_TABLES][B200_TPCH_MAX_COLS] = {
    /* lineitem */
    {{"l_orderkey", K_I64}, {"l_partkey", K_I64}, {"l_suppkey", K_I64}, {"l_linenumber", K_I32},
     {"l_quantity", K_DEC}, {"l_extendedprice", K_DEC}, {"l_discount", K_DEC}, {"l_tax", K_DEC},
     {"l_returnflag", K_STR}, {"l_linestatus", K_STR}, {"l_shipdate", K_DATE}, {"l_commitdate", K_DATE},
     {"l_receiptdate", K_DATE}, {"l_shipinstruct", K_STR}, {"l_shipmode", K_STR}, {"l_comment", K_STR}},
    /* orders */
    {{"o_orderkey", K_I64}, {"o_custkey", K_I64}, {"o_orderstatus", K_STR}, {"o_totalprice", K_DEC},
     {"o_orderdate", K_DATE}, {"o_orderpriority", K_STR}, {"o_clerk", K_STR}, {"o_shippriority", K_I32},
     {"o_comment", K_STR}},
    /* customer */
    {{"c_custkey", K_I64}, {"c_name", K_STR}, {"c_address", K_STR}, {"c_nationkey", K_I64},
     {"c_phone", K_STR}, {"c_acctbal", K_DEC}, {"c_mktsegment", K_STR}, {"c_comment", K_STR}},
    /* supplier */
    {{"s_suppkey", K_I64}, {"s_name", K_STR}, {"s_address", K_STR}, {"s_nationkey", K_I64},
     {"s_phone", K_STR}, {"s_acctbal", K_DEC}, {"s_comment", K_STR}},
    /* part */
    {{"p_partkey", K_I64}, {"p_name", K_STR}, {"p_mfgr", K_STR}, {"p_brand", K_STR}, {"p_type", K_STR},
     {"p_size", K_I32}, {"p_container", K_STR}, {"p_retailprice", K_DEC}, {"p_comment", K_STR}},
    /* partsupp */
    {{"ps_partkey", K_I64}, {"ps_suppkey", K_I64}, {"ps_availqty", K_I32}, {"ps_supplycost", K_DEC},
     {"ps_comment", K_STR}},
    /* nation */
    {{"n_nationkey", K_I64}, {"n_name", K_STR}, {"n_regionkey", K_I64}, {"n_comment", K_STR}},
    /* region */
    {{"r_regionkey", K_I64}, {"r_name", K_STR}, {"r_comment", K_STR}},
};

static const int kMaxStrLen = 48;  // upper bound of any generated string (bytes)

// dates as days since 1970-01-01
static const int32_t kStartDate = 8035;    // 1992-01-01
static const int32_t kEndDate = 10591;     // 1998-12-31
static const int32_t kCurrentDate = 9298;  // 1995-06-17

// scale factor is passed as milli-SF (SF 1 == 1000) so that tests can use SF 0.001 .. 0.1
B200_HD int64_t n_suppliers(int64_t msf) { int64_t n = 10000 * msf / 1000; return n < 4 ? 4 : n; }
B200_HD int64_t n_parts(int64_t msf) { int64_t n = 200000 * msf / 1000; return n < 8 ? 8 : n; }
B200_HD int64_t n_customers(int64_t msf) { int64_t n = 150000 * msf / 1000; return n < 6 ? 6 : n; }
B200_HD int64_t n_orders(int64_t msf) { int64_t n = 1500000 * msf / 1000; return n < 7 ? 7 : n; }

// order o (0-based) owns 1 + (o % 7) lineitems (TPC-H: uniform 1..7); closed-form prefix
B200_HD int64_t lines_before_order(int64_t o) {
  int64_t c = o / 7, k = o % 7;
  return c * 28 + k * (k + 1) / 2;
}
B200_HD void line_to_order(int64_t j, int64_t* o, int32_t* linenumber) {
  int64_t c = j / 28;
  int32_t rem = (int32_t)(j % 28);
  int32_t k = 0;
  while ((k + 1) * (k + 2) / 2 <= rem) k++;
  *o = c * 7 + k;
  *linenumber = rem - k * (k + 1) / 2 + 1;
}

B200_HD int64_t table_rows(int t, int64_t msf) {
  switch (t) {
    case LINEITEM: {
      int64_t total = lines_before_order(n_orders(msf));
      // spec row counts where the cycle scheme can reach them (SF10: 59,986,052)
      if (msf == 10000 && total >= 59986052) return 59986052;
      return total;
    }
    case ORDERS: return n_orders(msf);
    case CUSTOMER: return n_customers(msf);
    case SUPPLIER: return n_suppliers(msf);
    case PART: return n_parts(msf);
    case PARTSUPP: return n_parts(msf) * 4;
    case NATION: return 25;
    case REGION: return 5;
  }
  return 0;
}

B200_HD uint64_t rnd(int t, int col, int64_t row) {
  return mix64(mix64((uint64_t)row + 0x9E3779B97F4A7C15ull * (uint64_t)(col + 1)) ^ (0xB2000000ull + (uint64_t)t));
}

B200_HD int64_t order_key(int64_t o) { return (o / 8) * 32 + (o % 8) + 1; }  // sparse keys, first 8 of every 32
B200_HD int32_t order_date(int64_t o) { return kStartDate + (int32_t)(rnd(ORDERS, 4, o) % (uint64_t)(kEndDate - 151 - kStartDate + 1)); }
B200_HD int64_t retail_price_cents(int64_t partkey) { return 90000 + ((partkey / 10) % 20001) + 100 * (partkey % 1000); }
B200_HD int64_t supp_for_part(int64_t partkey, int i, int64_t msf) {
  int64_t S = n_suppliers(msf);
  return (partkey + (int64_t)i * (S / 4 + (partkey - 1) / S)) % S + 1;
}

struct LineDerived {
  int64_t o;
  int32_t ln;
  int64_t partkey;
  int32_t odate, ship, commit, receipt;
  int64_t qty;
};
B200_HD LineDerived line_derive(int64_t j, int64_t msf) {
  LineDerived d;
  line_to_order(j, &d.o, &d.ln);
  d.partkey = 1 + (int64_t)(rnd(LINEITEM, 1, j) % (uint64_t)n_parts(msf));
  d.odate = order_date(d.o);
  d.ship = d.odate + 1 + (int32_t)(rnd(LINEITEM, 10, j) % 121);
  d.commit = d.odate + 30 + (int32_t)(rnd(LINEITEM, 11, j) % 61);
  d.receipt = d.ship + 1 + (int32_t)(rnd(LINEITEM, 12, j) % 30);
  d.qty = 1 + (int64_t)(rnd(LINEITEM, 4, j) % 50);
  return d;
}

// Fixed-width columns: value as int64 (decimals: unscaled, scale 2; dates: days)
B200_HD int64_t gen_i64(int t, int col, int64_t row, int64_t msf) {
  switch (t) {
    case LINEITEM: {
      LineDerived d = line_derive(row, msf);
      switch (col) {
        case 0: return order_key(d.o);
        case 1: return d.partkey;
        case 2: return supp_for_part(d.partkey, (int)(rnd(LINEITEM, 2, row) % 4), msf);
        case 3: return d.ln;
        case 4: return d.qty * 100;
        case 5: return d.qty * retail_price_cents(d.partkey);
        case 6: return (int64_t)(rnd(LINEITEM, 6, row) % 11);
        case 7: return (int64_t)(rnd(LINEITEM, 7, row) % 9);
        case 10: return d.ship;
        case 11: return d.commit;
        case 12: return d.receipt;
      }
      return 0;
    }
    case ORDERS:
      switch (col) {
        case 0: return order_key(row);
        case 1: {  // customer keys never divisible by 3 (TPC-H 4.2.3)
          int64_t C = n_customers(msf);
          int64_t third = C / 3 > 0 ? C / 3 : 1;
          uint64_t h = rnd(ORDERS, 1, row);
          int64_t ck = (int64_t)(h % (uint64_t)third) * 3 + 1 + (int64_t)((h >> 40) & 1);
          return ck > C ? C - (C % 3 == 0 ? 1 : 0) : ck;
        }
        case 3: return 100000 + (int64_t)(rnd(ORDERS, 3, row) % 50000000);
        case 4: return order_date(row);
        case 7: return 0;
      }
      return 0;
    case CUSTOMER:
      switch (col) {
        case 0: return row + 1;
        case 3: return (int64_t)(rnd(CUSTOMER, 3, row) % 25);
        case 5: return -99999 + (int64_t)(rnd(CUSTOMER, 5, row) % 1099999);
      }
      return 0;
    case SUPPLIER:
      switch (col) {
        case 0: return row + 1;
        case 3: return (int64_t)(rnd(SUPPLIER, 3, row) % 25);
        case 5: return -99999 + (int64_t)(rnd(SUPPLIER, 5, row) % 1099999);
      }
      return 0;
    case PART:
      switch (col) {
        case 0: return row + 1;
        case 5: return 1 + (int64_t)(rnd(PART, 5, row) % 50);
        case 7: return retail_price_cents(row + 1);
      }
      return 0;
    case PARTSUPP:
      switch (col) {
        case 0: return row / 4 + 1;
        case 1: return supp_for_part(row / 4 + 1, (int)(row % 4), msf);
        case 2: return 1 + (int64_t)(rnd(PARTSUPP, 2, row) % 9999);
        case 3: return 100 + (int64_t)(rnd(PARTSUPP, 3, row) % 99901);
      }
      return 0;
    case NATION: {
      const int8_t region_of[25] = {0, 1, 1, 1, 4, 0, 3, 3, 2, 2, 4, 4, 2, 4, 0, 0, 0, 1, 2, 3, 4, 2, 3, 3, 1};
      switch (col) {
        case 0: return row;
        case 2: return region_of[row % 25];
      }
      return 0;
    }
    case REGION: return row;
  }
  return 0;
}

B200_HD uint32_t put(char* out, const char* s) {
  uint32_t n = 0;
  while (s[n]) {
    out[n] = s[n];
    n++;
  }
  return n;
}
B200_HD uint32_t put_num(char* out, uint64_t v, int width) {
  for (int i = width - 1; i >= 0; i--) {
    out[i] = (char)('0' + v % 10);
    v /= 10;
  }
  return (uint32_t)width;
}
B200_HD uint32_t put_word(char* out, uint64_t h, int minlen, int span) {
  uint32_t n = (uint32_t)(minlen + (int)(h % (uint64_t)span));
  uint64_t x = h;
  for (uint32_t i = 0; i < n; i++) {
    x = mix64(x + i);
    out[i] = (char)('a' + x % 26);
  }
  return n;
}

// String columns. `out` must hold kMaxStrLen bytes. Returns the byte length.
B200_HD uint32_t gen_str(int t, int col, int64_t row, int64_t msf, char* out) {
  const char* const kInstruct[4] = {"DELIVER IN PERSON", "COLLECT COD", "NONE", "TAKE BACK RETURN"};
  const char* const kModes[7] = {"REG AIR", "AIR", "RAIL", "SHIP", "TRUCK", "MAIL", "FOB"};
  const char* const kPrio[5] = {"1-URGENT", "2-HIGH", "3-MEDIUM", "4-NOT SPECIFIED", "5-LOW"};
  const char* const kSegments[5] = {"AUTOMOBILE", "BUILDING", "FURNITURE", "MACHINERY", "HOUSEHOLD"};
  const char* const kNations[25] = {"ALGERIA", "ARGENTINA", "BRAZIL", "CANADA", "EGYPT", "ETHIOPIA", "FRANCE",
                                    "GERMANY", "INDIA", "INDONESIA", "IRAN", "IRAQ", "JAPAN", "JORDAN", "KENYA",
                                    "MOROCCO", "MOZAMBIQUE", "PERU", "CHINA", "ROMANIA", "SAUDI ARABIA", "VIETNAM",
                                    "RUSSIA", "UNITED KINGDOM", "UNITED STATES"};
  const char* const kRegions[5] = {"AFRICA", "AMERICA", "ASIA", "EUROPE", "MIDDLE EAST"};
  const char* const kTypeA[6] = {"STANDARD", "SMALL", "MEDIUM", "LARGE", "ECONOMY", "PROMO"};
  const char* const kTypeB[5] = {"ANODIZED", "BURNISHED", "PLATED", "POLISHED", "BRUSHED"};
  const char* const kTypeC[5] = {"TIN", "NICKEL", "BRASS", "STEEL", "COPPER"};
  const char* const kContA[5] = {"SM", "LG", "MED", "JUMBO", "WRAP"};
  const char* const kContB[8] = {"CASE", "BOX", "BAG", "JAR", "PKG", "PACK", "CAN", "DRUM"};
  const char* const kColors[16] = {"almond", "antique", "aquamarine", "azure", "beige", "bisque", "black", "blanched",
                                   "blue", "blush", "brown", "burlywood", "chartreuse", "forest", "green", "khaki"};
  uint64_t h = rnd(t, col, row);
  uint32_t n = 0;
  switch (t) {
    case LINEITEM: {
      if (col == 8 || col == 9) {
        LineDerived d = line_derive(row, msf);
        if (col == 8) out[0] = d.receipt <= kCurrentDate ? ((h & 1) ? 'R' : 'A') : 'N';
        else out[0] = d.ship > kCurrentDate ? 'O' : 'F';
        return 1;
      }
      if (col == 13) return put(out, kInstruct[h % 4]);
      if (col == 14) return put(out, kModes[h % 7]);
      return put_word(out, h, 10, 34);
    }
    case ORDERS: {
      if (col == 2) {  // status derived from the order's lines: all shipped -> F, none -> O, else P
        int64_t j0 = lines_before_order(row);
        int nl = 1 + (int)(row % 7), nO = 0;
        for (int k = 0; k < nl; k++) nO += line_derive(j0 + k, msf).ship > kCurrentDate;
        out[0] = nO == nl ? 'O' : (nO == 0 ? 'F' : 'P');
        return 1;
      }
      if (col == 5) return put(out, kPrio[h % 5]);
      if (col == 6) {
        n = put(out, "Clerk#");
        int64_t nclerk = 1000 * msf / 1000;
        if (nclerk < 1) nclerk = 1;
        return n + put_num(out + n, 1 + h % (uint64_t)nclerk, 9);
      }
      return put_word(out, h, 19, 29);
    }
    case CUSTOMER: {
      if (col == 1) {
        n = put(out, "Customer#");
        return n + put_num(out + n, (uint64_t)row + 1, 9);
      }
      if (col == 4) {  // phone: CC-LLL-LLL-LLLL, CC = nationkey + 10
        int64_t nk = gen_i64(CUSTOMER, 3, row, msf);
        n = put_num(out, (uint64_t)nk + 10, 2);
        out[n++] = '-';
        n += put_num(out + n, 100 + (h >> 8) % 900, 3);
        out[n++] = '-';
        n += put_num(out + n, 100 + (h >> 20) % 900, 3);
        out[n++] = '-';
        n += put_num(out + n, 1000 + (h >> 32) % 9000, 4);
        return n;
      }
      if (col == 6) return put(out, kSegments[h % 5]);
      return put_word(out, h, 10, 30);
    }
    case SUPPLIER: {
      if (col == 1) {
        n = put(out, "Supplier#");
        return n + put_num(out + n, (uint64_t)row + 1, 9);
      }
      if (col == 4) {
        int64_t nk = gen_i64(SUPPLIER, 3, row, msf);
        n = put_num(out, (uint64_t)nk + 10, 2);
        out[n++] = '-';
        n += put_num(out + n, 100 + (h >> 8) % 900, 3);
        out[n++] = '-';
        n += put_num(out + n, 100 + (h >> 20) % 900, 3);
        out[n++] = '-';
        n += put_num(out + n, 1000 + (h >> 32) % 9000, 4);
        return n;
      }
      if (col == 6) {  // ~0.05% "Customer ... Complaints" (q16)
        if (h % 2000 == 0) return put(out, "sly Customer bold Complaints wake");
        return put_word(out, h, 25, 20);
      }
      return put_word(out, h, 10, 30);
    }
    case PART: {
      uint64_t hm = rnd(PART, 2, row);
      int m = 1 + (int)(hm % 5);
      if (col == 1) {
        n = put(out, kColors[h % 16]);
        out[n++] = ' ';
        n += put(out + n, kColors[(h >> 8) % 16]);
        out[n++] = ' ';
        n += put(out + n, kColors[(h >> 16) % 16]);
        return n;
      }
      if (col == 2) {
        n = put(out, "Manufacturer#");
        out[n++] = (char)('0' + m);
        return n;
      }
      if (col == 3) {
        n = put(out, "Brand#");
        out[n++] = (char)('0' + m);
        out[n++] = (char)('1' + (int)(rnd(PART, 3, row) % 5));
        return n;
      }
      if (col == 4) {
        n = put(out, kTypeA[h % 6]);
        out[n++] = ' ';
        n += put(out + n, kTypeB[(h >> 8) % 5]);
        out[n++] = ' ';
        n += put(out + n, kTypeC[(h >> 16) % 5]);
        return n;
      }
      if (col == 6) {
        n = put(out, kContA[h % 5]);
        out[n++] = ' ';
        n += put(out + n, kContB[(h >> 8) % 8]);
        return n;
      }
      return put_word(out, h, 5, 18);
    }
    case PARTSUPP: return put_word(out, h, 20, 28);
    case NATION:
      if (col == 1) return put(out, kNations[row % 25]);
      return put_word(out, h, 20, 28);
    case REGION:
      if (col == 1) return put(out, kRegions[row % 5]);
      return put_word(out, h, 20, 28);
  }
  return 0;
}

}  // namespace tpch
}  // namespace b200
