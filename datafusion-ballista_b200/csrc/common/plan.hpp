// Stage-plan IR: the physical operator tree a Ballista task hands to an ExecutionEngine
// (ballista/executor/src/execution_engine.rs:50-58), restated as a small typed tree parsed from
// JSON.  Node and field names follow the vendored DataFusion plan protobuf that pins the shape of
// every operator the executor can receive: ballista/core/proto/datafusion.proto
//   FilterExecNode :1027-1034, ProjectionExecNode :1211-1215, AggregateExecNode :1257-1271
//   (modes :1217-1224), HashJoinExecNode :1134-1144 (PartitionMode :1128-1132), SortExecNode
//   :1286-1292, SortPreservingMergeExecNode :1294-1298, PhysicalExprNode :851-901,
//   PhysicalBinaryExprNode :957-961; ShuffleWriterExecNode / ShuffleReaderExecNode are Ballista's
//   own (ballista/core/proto/ballista.proto:47-99).
//
// Type rules marked [EXT] restate arrow-rs 58.1 / DataFusion 53.1 behaviour (crates not vendored
// under /root/reference, Cargo.lock:192,2053) and are documented in DESIGN.md §"Semantics".
//
// This header is shared by the product (csrc/host) and by the CPU oracle (oracle/): it contains
// parsing and *typing* only -- no arithmetic on data.
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "json.hpp"

namespace b200 {

typedef __int128 i128;
typedef unsigned __int128 u128;

enum class TypeId : uint8_t {
  Null = 0, Bool, Int8, Int16, Int32, Int64, UInt8, UInt16, UInt32, UInt64,
  Float32, Float64, Date32, Timestamp, Decimal128, Utf8
};

// Physical (compute) kind a logical type maps onto inside both engines.
enum class PK : uint8_t { Bool = 0, I64 = 1, F64 = 2, I128 = 3, Str = 4 };

struct DataType {
  TypeId id = TypeId::Null;
  uint8_t precision = 0;
  int8_t scale = 0;
  DataType() {}
  DataType(TypeId i) : id(i) {}
  static DataType decimal(int p, int s) {
    DataType t(TypeId::Decimal128);
    t.precision = (uint8_t)p;
    t.scale = (int8_t)s;
    return t;
  }
  bool operator==(const DataType& o) const {
    return id == o.id && (id != TypeId::Decimal128 || (precision == o.precision && scale == o.scale));
  }
  bool operator!=(const DataType& o) const { return !(*this == o); }
  bool is_decimal() const { return id == TypeId::Decimal128; }
  bool is_float() const { return id == TypeId::Float32 || id == TypeId::Float64; }
  bool is_signed_int() const { return id >= TypeId::Int8 && id <= TypeId::Int64; }
  bool is_unsigned_int() const { return id >= TypeId::UInt8 && id <= TypeId::UInt64; }
  bool is_integer() const { return is_signed_int() || is_unsigned_int(); }
  bool is_numeric() const { return is_integer() || is_float() || is_decimal(); }
  bool is_string() const { return id == TypeId::Utf8; }
  PK pk() const {
    switch (id) {
      case TypeId::Bool: return PK::Bool;
      case TypeId::Float32:
      case TypeId::Float64: return PK::F64;
      case TypeId::Decimal128: return PK::I128;
      case TypeId::Utf8: return PK::Str;
      default: return PK::I64;
    }
  }
  // Arrow in-memory width of one value in the values buffer (Utf8: the int32 offset).
  int width() const {
    switch (id) {
      case TypeId::Null: return 0;
      case TypeId::Bool: return 0;  // bit-packed
      case TypeId::Int8:
      case TypeId::UInt8: return 1;
      case TypeId::Int16:
      case TypeId::UInt16: return 2;
      case TypeId::Int32:
      case TypeId::UInt32:
      case TypeId::Float32:
      case TypeId::Date32: return 4;
      case TypeId::Decimal128: return 16;
      case TypeId::Utf8: return 4;
      default: return 8;
    }
  }
  std::string str() const {
    switch (id) {
      case TypeId::Null: return "null";
      case TypeId::Bool: return "bool";
      case TypeId::Int8: return "i8";
      case TypeId::Int16: return "i16";
      case TypeId::Int32: return "i32";
      case TypeId::Int64: return "i64";
      case TypeId::UInt8: return "u8";
      case TypeId::UInt16: return "u16";
      case TypeId::UInt32: return "u32";
      case TypeId::UInt64: return "u64";
      case TypeId::Float32: return "f32";
      case TypeId::Float64: return "f64";
      case TypeId::Date32: return "date32";
      case TypeId::Timestamp: return "ts";
      case TypeId::Utf8: return "utf8";
      case TypeId::Decimal128:
        return "dec(" + std::to_string((int)precision) + "," + std::to_string((int)scale) + ")";
    }
    return "?";
  }
};

inline DataType parse_type(const Json& j) {
  if (j.is_obj()) {
    const Json& d = j.at("dec");
    return DataType::decimal((int)d.at(0).as_int(), (int)d.at(1).as_int());
  }
  const std::string& s = j.str();
  static const std::pair<const char*, TypeId> tab[] = {
      {"null", TypeId::Null},     {"bool", TypeId::Bool},     {"i8", TypeId::Int8},
      {"i16", TypeId::Int16},     {"i32", TypeId::Int32},     {"i64", TypeId::Int64},
      {"u8", TypeId::UInt8},      {"u16", TypeId::UInt16},    {"u32", TypeId::UInt32},
      {"u64", TypeId::UInt64},    {"f32", TypeId::Float32},   {"f64", TypeId::Float64},
      {"date32", TypeId::Date32}, {"ts", TypeId::Timestamp},  {"utf8", TypeId::Utf8}};
  for (auto& kv : tab)
    if (s == kv.first) return DataType(kv.second);
  throw std::runtime_error("plan IR: unknown type '" + s + "'");
}

inline std::string type_json(const DataType& t) {
  if (t.is_decimal())
    return "{\"dec\":[" + std::to_string((int)t.precision) + "," + std::to_string((int)t.scale) + "]}";
  return "\"" + t.str() + "\"";
}

struct Field {
  std::string name;
  DataType type;
  bool nullable = true;
};
typedef std::vector<Field> Schema;

inline Schema parse_schema(const Json& j) {
  Schema s;
  for (size_t i = 0; i < j.size(); i++) {
    const Json& f = j.at(i);
    Field fd;
    fd.name = f.at("name").str();
    fd.type = parse_type(f.at("type"));
    fd.nullable = f.get_bool("nullable", true);
    s.push_back(fd);
  }
  return s;
}

// ----------------------------------------------------------------------------------------------
// Decimal helpers (typing only)
// ----------------------------------------------------------------------------------------------
static const int kMaxDecimalPrecision = 38;
static const int kMaxDecimalScale = 38;

inline i128 pow10_i128(int n) {
  i128 r = 1;
  for (int i = 0; i < n; i++) r *= 10;
  return r;
}

inline i128 parse_i128(const std::string& s) {
  size_t k = 0;
  bool neg = false;
  if (k < s.size() && (s[k] == '-' || s[k] == '+')) neg = s[k++] == '-';
  u128 v = 0;
  if (k >= s.size()) throw std::runtime_error("plan IR: bad integer literal '" + s + "'");
  for (; k < s.size(); k++) {
    if (s[k] < '0' || s[k] > '9') throw std::runtime_error("plan IR: bad integer literal '" + s + "'");
    v = v * 10 + (unsigned)(s[k] - '0');
  }
  return neg ? -(i128)v : (i128)v;
}

inline std::string i128_to_string(i128 v) {
  if (v == 0) return "0";
  bool neg = v < 0;
  u128 u = neg ? (u128)(-(v + 1)) + 1 : (u128)v;
  std::string s;
  while (u) {
    s += (char)('0' + (int)(u % 10));
    u /= 10;
  }
  if (neg) s += '-';
  return std::string(s.rbegin(), s.rend());
}

// Integer -> decimal coercion precision used by DataFusion when an integer meets a decimal
// (datafusion-expr type_coercion/binary.rs `coerce_numeric_type_to_decimal`) [EXT].
inline DataType int_as_decimal(const DataType& t) {
  switch (t.id) {
    case TypeId::Int8: return DataType::decimal(3, 0);
    case TypeId::Int16: return DataType::decimal(5, 0);
    case TypeId::Int32: return DataType::decimal(10, 0);
    case TypeId::Int64: return DataType::decimal(20, 0);
    case TypeId::UInt8: return DataType::decimal(3, 0);
    case TypeId::UInt16: return DataType::decimal(5, 0);
    case TypeId::UInt32: return DataType::decimal(10, 0);
    case TypeId::UInt64: return DataType::decimal(20, 0);
    default: return t;
  }
}

enum class BinOp : uint8_t { Add, Sub, Mul, Div, Mod, Eq, Ne, Lt, Le, Gt, Ge, And, Or };

inline BinOp parse_binop(const std::string& s) {
  static const std::pair<const char*, BinOp> tab[] = {
      {"+", BinOp::Add},  {"-", BinOp::Sub},  {"*", BinOp::Mul},  {"/", BinOp::Div},
      {"%", BinOp::Mod},  {"=", BinOp::Eq},   {"==", BinOp::Eq},  {"!=", BinOp::Ne},
      {"<>", BinOp::Ne},  {"<", BinOp::Lt},   {"<=", BinOp::Le},  {">", BinOp::Gt},
      {">=", BinOp::Ge},  {"and", BinOp::And}, {"or", BinOp::Or},  {"AND", BinOp::And},
      {"OR", BinOp::Or}};
  for (auto& kv : tab)
    if (s == kv.first) return kv.second;
  throw std::runtime_error("plan IR: unknown binary operator '" + s + "'");
}
inline bool is_arith(BinOp o) { return o <= BinOp::Mod; }
inline bool is_compare(BinOp o) { return o >= BinOp::Eq && o <= BinOp::Ge; }
inline bool is_logic(BinOp o) { return o == BinOp::And || o == BinOp::Or; }

// Result type of decimal (op) decimal, following arrow-arith 58 `decimal_op` [EXT]:
//   add/sub: scale = max(s1,s2); precision = min(38, max(p1-s1,p2-s2) + scale + 1)
//   mul:     scale = s1+s2;      precision = min(38, p1+p2+1)
//   div:     scale = min(38, s1+4); precision = min(38, p1 + (scale - s1 + s2))
//   mod:     scale = max(s1,s2); precision = min(38, min(p1-s1,p2-s2) + scale)
inline DataType decimal_result_type(BinOp op, const DataType& a, const DataType& b) {
  int p1 = a.precision, s1 = a.scale, p2 = b.precision, s2 = b.scale;
  int p, s;
  switch (op) {
    case BinOp::Add:
    case BinOp::Sub:
      s = std::max(s1, s2);
      p = std::min(kMaxDecimalPrecision, std::max(p1 - s1, p2 - s2) + s + 1);
      break;
    case BinOp::Mul:
      s = s1 + s2;
      if (s > kMaxDecimalScale) throw std::runtime_error("decimal multiply: result scale exceeds 38");
      p = std::min(kMaxDecimalPrecision, p1 + p2 + 1);
      break;
    case BinOp::Div: {
      s = std::min(kMaxDecimalScale, s1 + 4);
      int mul_pow = s - s1 + s2;
      p = std::min(kMaxDecimalPrecision, p1 + mul_pow);
      break;
    }
    case BinOp::Mod:
      s = std::max(s1, s2);
      p = std::min(kMaxDecimalPrecision, std::min(p1 - s1, p2 - s2) + s);
      break;
    default: throw std::runtime_error("decimal_result_type: not arithmetic");
  }
  return DataType::decimal(p, s);
}

// ----------------------------------------------------------------------------------------------
// Expressions
// ----------------------------------------------------------------------------------------------
struct LitValue {
  bool is_null = false;
  int64_t i = 0;   // Bool / ints / Date32 / Timestamp
  double f = 0;    // floats
  i128 d = 0;      // Decimal128 unscaled
  std::string s;   // Utf8
};

struct Expr;
typedef std::shared_ptr<Expr> ExprPtr;

struct Expr {
  enum Kind { Col, Lit, Bin, Not, Neg, IsNull, IsNotNull, Cast, Case, InList, Like, Fn } kind = Lit;
  DataType type;          // resolved output type
  bool nullable = true;
  int col = -1;           // Col
  LitValue lit;           // Lit
  BinOp op = BinOp::Add;  // Bin
  std::string fn;         // Fn: "date_part_year", "substr"
  std::vector<ExprPtr> args;  // Bin: l,r; unary: x; Case: [w0,t0,w1,t1,...,(else)]; InList: x, items...; Fn args
  bool has_else = false;  // Case
  bool negated = false;   // InList / Like
  std::string pattern;    // Like
  std::string name;       // display only
};

inline ExprPtr make_col(int idx, const Schema& in) {
  if (idx < 0 || (size_t)idx >= in.size()) throw std::runtime_error("plan IR: column index out of range");
  auto e = std::make_shared<Expr>();
  e->kind = Expr::Col;
  e->col = idx;
  e->type = in[idx].type;
  e->nullable = in[idx].nullable;
  e->name = in[idx].name;
  return e;
}

// rewrite every column reference k -> target[k] (in place)
inline void remap_columns(const ExprPtr& e, const std::vector<int>& target) {
  if (!e) return;
  if (e->kind == Expr::Col) {
    if (e->col < 0 || (size_t)e->col >= target.size()) throw std::runtime_error("plan IR: column index out of range in remap");
    e->col = target[(size_t)e->col];
  }
  for (auto& a : e->args) remap_columns(a, target);
}

inline LitValue parse_lit_value(const DataType& t, const Json& v) {
  LitValue l;
  if (v.is_null()) {
    l.is_null = true;
    return l;
  }
  switch (t.pk()) {
    case PK::Bool: l.i = v.as_bool(); break;
    case PK::I64: l.i = v.is_str() ? (int64_t)parse_i128(v.str()) : v.as_int(); break;
    case PK::F64: l.f = v.is_str() ? strtod(v.str().c_str(), nullptr) : v.as_double(); break;
    case PK::I128: l.d = v.is_str() ? parse_i128(v.str()) : (v.is_int ? (i128)v.i : parse_i128(v.s)); break;
    case PK::Str: l.s = v.str(); break;
  }
  return l;
}

inline DataType arith_result_type(BinOp op, DataType a, DataType b) {
  if (a.id == TypeId::Null) return b;
  if (b.id == TypeId::Null) return a;
  if (a.is_float() || b.is_float()) {
    if (a.id == TypeId::Float32 && b.id == TypeId::Float32) return DataType(TypeId::Float32);
    return DataType(TypeId::Float64);
  }
  if (a.is_decimal() || b.is_decimal()) {
    if (!a.is_decimal()) a = int_as_decimal(a);
    if (!b.is_decimal()) b = int_as_decimal(b);
    if (!a.is_decimal() || !b.is_decimal())
      throw std::runtime_error("arithmetic between " + a.str() + " and " + b.str() + " is not supported");
    return decimal_result_type(op, a, b);
  }
  if (a.is_integer() && b.is_integer()) {
    if (a == b) return a;
    return DataType(TypeId::Int64);
  }
  if (a.id == TypeId::Date32 && b.is_integer() && (op == BinOp::Add || op == BinOp::Sub)) return a;
  if (a.id == TypeId::Date32 && b.id == TypeId::Date32 && op == BinOp::Sub) return DataType(TypeId::Int64);
  throw std::runtime_error("arithmetic between " + a.str() + " and " + b.str() + " is not supported");
}

ExprPtr parse_expr(const Json& j, const Schema& in);

inline ExprPtr parse_expr(const Json& j, const Schema& in) {
  auto e = std::make_shared<Expr>();
  if (j.has("col") || j.find("col")) {
    const Json& c = j.at("col");
    if (c.is_str()) {
      for (size_t i = 0; i < in.size(); i++)
        if (in[i].name == c.str()) return make_col((int)i, in);
      throw std::runtime_error("plan IR: unknown column '" + c.str() + "'");
    }
    return make_col((int)c.as_int(), in);
  }
  if (j.find("lit")) {
    const Json& l = j.at("lit");
    e->kind = Expr::Lit;
    e->type = parse_type(l.at("t"));
    const Json* v = l.find("v");
    Json nullj;
    e->lit = parse_lit_value(e->type, v ? *v : nullj);
    e->nullable = e->lit.is_null;
    return e;
  }
  if (j.find("bin")) {
    e->kind = Expr::Bin;
    e->op = parse_binop(j.at("bin").str());
    e->args.push_back(parse_expr(j.at("l"), in));
    e->args.push_back(parse_expr(j.at("r"), in));
    const DataType& a = e->args[0]->type;
    const DataType& b = e->args[1]->type;
    if (is_arith(e->op)) e->type = arith_result_type(e->op, a, b);
    else e->type = DataType(TypeId::Bool);
    if (is_logic(e->op) && (a.id != TypeId::Bool || b.id != TypeId::Bool) && a.id != TypeId::Null && b.id != TypeId::Null)
      throw std::runtime_error("AND/OR need boolean operands");
    if (is_compare(e->op)) {
      bool ok = (a.pk() == b.pk()) || (a.is_numeric() && b.is_numeric()) || a.id == TypeId::Null || b.id == TypeId::Null;
      if (!ok) throw std::runtime_error("cannot compare " + a.str() + " with " + b.str());
    }
    e->nullable = e->args[0]->nullable || e->args[1]->nullable || e->op == BinOp::Div || e->op == BinOp::Mod;
    return e;
  }
  if (j.find("not")) {
    e->kind = Expr::Not;
    e->args.push_back(parse_expr(j.at("not"), in));
    e->type = DataType(TypeId::Bool);
    e->nullable = e->args[0]->nullable;
    return e;
  }
  if (j.find("neg")) {
    e->kind = Expr::Neg;
    e->args.push_back(parse_expr(j.at("neg"), in));
    e->type = e->args[0]->type;
    e->nullable = e->args[0]->nullable;
    return e;
  }
  if (j.find("is_null") || j.find("is_not_null")) {
    bool isn = j.find("is_null") != nullptr;
    e->kind = isn ? Expr::IsNull : Expr::IsNotNull;
    e->args.push_back(parse_expr(j.at(isn ? "is_null" : "is_not_null"), in));
    e->type = DataType(TypeId::Bool);
    e->nullable = false;
    return e;
  }
  if (j.find("cast")) {
    e->kind = Expr::Cast;
    e->args.push_back(parse_expr(j.at("cast"), in));
    e->type = parse_type(j.at("to"));
    e->nullable = e->args[0]->nullable;
    return e;
  }
  if (j.find("case")) {
    e->kind = Expr::Case;
    const Json& c = j.at("case");
    const Json& whens = c.at("when");
    DataType rt;
    for (size_t i = 0; i < whens.size(); i++) {
      e->args.push_back(parse_expr(whens.at(i).at(0), in));
      e->args.push_back(parse_expr(whens.at(i).at(1), in));
      if (rt.id == TypeId::Null) rt = e->args.back()->type;
    }
    if (c.has("else")) {
      e->args.push_back(parse_expr(c.at("else"), in));
      e->has_else = true;
      if (rt.id == TypeId::Null) rt = e->args.back()->type;
    }
    e->type = rt;
    e->nullable = true;
    return e;
  }
  if (j.find("in")) {
    e->kind = Expr::InList;
    e->args.push_back(parse_expr(j.at("in"), in));
    const Json& lst = j.at("list");
    for (size_t i = 0; i < lst.size(); i++) e->args.push_back(parse_expr(lst.at(i), in));
    e->negated = j.get_bool("negated", false);
    e->type = DataType(TypeId::Bool);
    e->nullable = e->args[0]->nullable;
    return e;
  }
  if (j.find("like")) {
    e->kind = Expr::Like;
    e->args.push_back(parse_expr(j.at("like"), in));
    e->pattern = j.at("pattern").str();
    e->negated = j.get_bool("negated", false);
    e->type = DataType(TypeId::Bool);
    e->nullable = e->args[0]->nullable;
    if (!e->args[0]->type.is_string()) throw std::runtime_error("LIKE needs a utf8 operand");
    return e;
  }
  if (j.find("fn")) {
    e->kind = Expr::Fn;
    e->fn = j.at("fn").str();
    const Json& as = j.at("args");
    for (size_t i = 0; i < as.size(); i++) e->args.push_back(parse_expr(as.at(i), in));
    if (e->fn == "date_part_year") {
      // DataFusion: date_part('year', Date32) -> Int32 [EXT]
      e->type = DataType(TypeId::Int32);
    } else if (e->fn == "substr") {
      e->type = DataType(TypeId::Utf8);
    } else {
      throw std::runtime_error("plan IR: unknown scalar function '" + e->fn + "'");
    }
    e->nullable = e->args.empty() ? false : e->args[0]->nullable;
    return e;
  }
  throw std::runtime_error("plan IR: unrecognised expression node");
}

// ----------------------------------------------------------------------------------------------
// Aggregates
// ----------------------------------------------------------------------------------------------
enum class AggFn : uint8_t { Sum, Min, Max, Count, Avg };
enum class AggMode : uint8_t { Partial, Final, FinalPartitioned, Single, SinglePartitioned };

inline bool agg_mode_consumes_states(AggMode m) { return m == AggMode::Final || m == AggMode::FinalPartitioned; }
inline bool agg_mode_emits_states(AggMode m) { return m == AggMode::Partial; }

struct AggExpr {
  AggFn fn = AggFn::Sum;
  ExprPtr arg;           // null for COUNT(*) and in Final modes
  DataType input_type;   // type of arg (after AVG's integer->f64 coercion); for Final: taken from IR
  DataType sum_type;     // accumulator type for Sum/Avg
  DataType result_type;  // final value type
  bool distinct = false;
  std::string name;
  int n_state_cols() const { return fn == AggFn::Avg ? 2 : 1; }
};

// [EXT] datafusion-functions-aggregate 53: SUM(Decimal128(p,s)) -> Decimal128(min(38,p+10), s);
// SUM(int) -> Int64, SUM(uint) -> UInt64, SUM(float) -> Float64.
inline DataType sum_result_type(const DataType& t) {
  if (t.is_decimal()) return DataType::decimal(std::min(kMaxDecimalPrecision, t.precision + 10), t.scale);
  if (t.is_signed_int()) return DataType(TypeId::Int64);
  if (t.is_unsigned_int()) return DataType(TypeId::UInt64);
  if (t.is_float()) return DataType(TypeId::Float64);
  throw std::runtime_error("SUM does not support " + t.str());
}
// [EXT] AVG(Decimal128(p,s)) -> Decimal128(min(38,p+4), min(38,s+4)); AVG(other numeric) -> Float64
inline DataType avg_result_type(const DataType& t) {
  if (t.is_decimal())
    return DataType::decimal(std::min(kMaxDecimalPrecision, t.precision + 4), std::min(kMaxDecimalScale, t.scale + 4));
  if (t.is_numeric()) return DataType(TypeId::Float64);
  throw std::runtime_error("AVG does not support " + t.str());
}

inline AggFn parse_aggfn(const std::string& s) {
  if (s == "sum") return AggFn::Sum;
  if (s == "min") return AggFn::Min;
  if (s == "max") return AggFn::Max;
  if (s == "count") return AggFn::Count;
  if (s == "avg") return AggFn::Avg;
  throw std::runtime_error("plan IR: unknown aggregate '" + s + "'");
}
inline AggMode parse_aggmode(const std::string& s) {
  if (s == "Partial") return AggMode::Partial;
  if (s == "Final") return AggMode::Final;
  if (s == "FinalPartitioned") return AggMode::FinalPartitioned;
  if (s == "Single") return AggMode::Single;
  if (s == "SinglePartitioned") return AggMode::SinglePartitioned;
  throw std::runtime_error("plan IR: unknown aggregate mode '" + s + "'");
}

// ----------------------------------------------------------------------------------------------
// Plan nodes
// ----------------------------------------------------------------------------------------------
enum class JoinType : uint8_t { Inner, Left, Right, Full, LeftSemi, RightSemi, LeftAnti, RightAnti };
inline JoinType parse_join_type(const std::string& s) {
  static const std::pair<const char*, JoinType> tab[] = {
      {"Inner", JoinType::Inner},         {"Left", JoinType::Left},           {"Right", JoinType::Right},
      {"Full", JoinType::Full},           {"LeftSemi", JoinType::LeftSemi},   {"RightSemi", JoinType::RightSemi},
      {"LeftAnti", JoinType::LeftAnti},   {"RightAnti", JoinType::RightAnti}};
  for (auto& kv : tab)
    if (s == kv.first) return kv.second;
  throw std::runtime_error("plan IR: unknown join type '" + s + "'");
}

struct SortKey {
  ExprPtr expr;
  bool asc = true;
  bool nulls_first = false;
};

struct NamedExpr {
  ExprPtr expr;
  std::string name;
};

struct PlanNode;
typedef std::unique_ptr<PlanNode> PlanPtr;

struct PlanNode {
  enum Op {
    Scan, ShuffleReader, Filter, Projection, Aggregate, HashJoin, Sort, SortPreservingMerge,
    Passthrough /* CoalesceBatches, CoalescePartitions, round-robin Repartition */, Limit, ShuffleWriter
  } op = Scan;
  std::string op_name;
  std::vector<PlanPtr> children;
  Schema schema;  // output schema

  // Scan
  std::string table;
  std::vector<int> scan_projection;  // indices into the registered table's schema
  // ShuffleReader
  int64_t reader_stage_id = 0;
  bool broadcast = false;
  // Filter
  ExprPtr predicate;
  std::vector<int> projection;  // Filter / HashJoin optional output projection
  bool has_projection = false;
  // Projection
  std::vector<NamedExpr> exprs;
  // Aggregate
  AggMode agg_mode = AggMode::Single;
  std::vector<NamedExpr> group_by;
  std::vector<AggExpr> aggs;
  // HashJoin
  JoinType join_type = JoinType::Inner;
  std::string partition_mode;  // CollectLeft | Partitioned
  std::vector<std::pair<ExprPtr, ExprPtr>> on;
  bool null_equals_null = false;
  ExprPtr join_filter;  // over concat(left schema, right schema)
  // Sort / SPM / Limit
  std::vector<SortKey> sort_keys;
  int64_t fetch = -1;
  int64_t skip = 0;
  bool preserve_partitioning = false;
  // ShuffleWriter
  std::string job_id;
  int64_t stage_id = 0;
  std::vector<ExprPtr> part_exprs;
  int64_t n_out_partitions = 0;  // 0 => no repartitioning ("None" branch, shuffle_writer.rs:221-268)
  bool sort_shuffle = true;
};

// deep copy of an expression with every column reference moved by `delta` positions
inline ExprPtr shift_cols(const ExprPtr& e, int delta) {
  if (!e || delta == 0) return e;
  auto c = std::make_shared<Expr>(*e);
  if (c->kind == Expr::Col) c->col += delta;
  for (auto& a : c->args) a = shift_cols(a, delta);
  return c;
}

inline std::vector<SortKey> parse_sort_keys(const Json& j, const Schema& in) {
  std::vector<SortKey> ks;
  for (size_t i = 0; i < j.size(); i++) {
    SortKey k;
    k.expr = parse_expr(j.at(i).at("expr"), in);
    k.asc = j.at(i).get_bool("asc", true);
    k.nulls_first = j.at(i).get_bool("nulls_first", !k.asc);  // SQL default: NULLS LAST for ASC, FIRST for DESC
    ks.push_back(k);
  }
  return ks;
}

PlanPtr parse_plan(const Json& j);

inline PlanPtr parse_plan(const Json& j) {
  auto n = PlanPtr(new PlanNode());
  const std::string& op = j.at("op").str();
  n->op_name = op;
  auto parse_child = [&](const char* key) {
    n->children.push_back(parse_plan(j.at(key)));
    return n->children.back().get();
  };
  if (op == "DataSourceExec" || op == "MemoryScan" || op == "Scan") {
    n->op = PlanNode::Scan;
    n->table = j.at("table").str();
    Schema full = parse_schema(j.at("schema"));
    if (j.has("projection")) {
      const Json& p = j.at("projection");
      for (size_t i = 0; i < p.size(); i++) {
        int idx = (int)p.at(i).as_int();
        if (idx < 0 || (size_t)idx >= full.size()) throw std::runtime_error("scan projection out of range");
        n->scan_projection.push_back(idx);
        n->schema.push_back(full[idx]);
      }
    } else {
      for (size_t i = 0; i < full.size(); i++) n->scan_projection.push_back((int)i);
      n->schema = full;
    }
  } else if (op == "ShuffleReaderExec" || op == "UnresolvedShuffleExec") {
    n->op = PlanNode::ShuffleReader;
    n->reader_stage_id = j.at("stage_id").as_int();
    n->schema = parse_schema(j.at("schema"));
    n->broadcast = j.get_bool("broadcast", false);
  } else if (op == "FilterExec") {
    n->op = PlanNode::Filter;
    PlanNode* c = parse_child("input");
    n->predicate = parse_expr(j.at("predicate"), c->schema);
    if (n->predicate->type.id != TypeId::Bool) throw std::runtime_error("FilterExec predicate must be boolean");
    if (j.has("projection")) {
      n->has_projection = true;
      const Json& p = j.at("projection");
      for (size_t i = 0; i < p.size(); i++) {
        int idx = (int)p.at(i).as_int();
        if (idx < 0 || (size_t)idx >= c->schema.size()) throw std::runtime_error("filter projection out of range");
        n->projection.push_back(idx);
        n->schema.push_back(c->schema[idx]);
      }
    } else {
      n->schema = c->schema;
    }
    n->fetch = j.get_int("fetch", -1);
  } else if (op == "ProjectionExec") {
    n->op = PlanNode::Projection;
    PlanNode* c = parse_child("input");
    const Json& es = j.at("exprs");
    for (size_t i = 0; i < es.size(); i++) {
      NamedExpr ne;
      ne.expr = parse_expr(es.at(i).at("expr"), c->schema);
      ne.name = es.at(i).get_str("name", ne.expr->name.empty() ? ("c" + std::to_string(i)) : ne.expr->name);
      n->exprs.push_back(ne);
      Field f;
      f.name = ne.name;
      f.type = ne.expr->type;
      f.nullable = ne.expr->nullable;
      n->schema.push_back(f);
    }
  } else if (op == "AggregateExec") {
    n->op = PlanNode::Aggregate;
    PlanNode* c = parse_child("input");
    n->agg_mode = parse_aggmode(j.at("mode").str());
    bool from_states = agg_mode_consumes_states(n->agg_mode);
    const Json& gs = j.at("group_by");
    for (size_t i = 0; i < gs.size(); i++) {
      NamedExpr ne;
      if (from_states) ne.expr = make_col((int)i, c->schema);  // group keys are the leading columns of the partial output
      else ne.expr = parse_expr(gs.at(i).at("expr"), c->schema);
      ne.name = gs.at(i).get_str("name", ne.expr->name.empty() ? ("g" + std::to_string(i)) : ne.expr->name);
      n->group_by.push_back(ne);
      Field f;
      f.name = ne.name;
      f.type = ne.expr->type;
      f.nullable = ne.expr->nullable;
      n->schema.push_back(f);
    }
    const Json& as = j.at("aggr");
    size_t state_col = gs.size();
    for (size_t i = 0; i < as.size(); i++) {
      const Json& a = as.at(i);
      AggExpr ae;
      ae.fn = parse_aggfn(a.at("fn").str());
      ae.distinct = a.get_bool("distinct", false);
      if (ae.distinct) throw std::runtime_error("DISTINCT aggregates must be lowered to two-level aggregation by the planner");
      ae.name = a.get_str("name", a.at("fn").str() + "_" + std::to_string(i));
      if (from_states) {
        // states are read positionally: AVG -> (count:UInt64, sum), others -> one column
        if (ae.fn == AggFn::Avg) {
          if (state_col + 1 >= c->schema.size()) throw std::runtime_error("Final aggregate: missing AVG state columns");
          ae.sum_type = c->schema[state_col + 1].type;
          if (a.has("input_type")) {
            ae.input_type = parse_type(a.at("input_type"));
          } else if (j.has("input_schema") && a.has("args") && a.at("args").size() > 0) {
            // the protobuf form (AggregateExecNode.input_schema = 7 + the original argument expression): typed here
            ae.input_type = parse_expr(a.at("args").at(0), parse_schema(j.at("input_schema")))->type;
          } else {
            throw std::runtime_error("Final AVG needs \"input_type\" (or the original argument and \"input_schema\")");
          }
          ae.result_type = avg_result_type(ae.input_type);
        } else {
          if (state_col >= c->schema.size()) throw std::runtime_error("Final aggregate: missing state column");
          ae.input_type = c->schema[state_col].type;
          ae.sum_type = ae.input_type;
          ae.result_type = ae.input_type;
        }
        state_col += ae.n_state_cols();
      } else {
        if (a.has("args") && a.at("args").size() > 0) {
          ae.arg = parse_expr(a.at("args").at(0), c->schema);
          // COUNT(<non-null literal>) is COUNT(*)
          if (ae.fn == AggFn::Count && ae.arg->kind == Expr::Lit && !ae.arg->lit.is_null) ae.arg = nullptr;
        }
        if (!ae.arg && ae.fn != AggFn::Count) throw std::runtime_error("aggregate needs an argument");
        DataType it = ae.arg ? ae.arg->type : DataType(TypeId::Int64);
        switch (ae.fn) {
          case AggFn::Sum:
            ae.input_type = it;
            ae.sum_type = sum_result_type(it);
            ae.result_type = ae.sum_type;
            break;
          case AggFn::Avg:
            ae.input_type = it;
            // [EXT] AVG over non-decimal numerics is computed in Float64 (input cast to f64 first)
            ae.sum_type = it.is_decimal() ? sum_result_type(it) : DataType(TypeId::Float64);
            ae.result_type = avg_result_type(it);
            break;
          case AggFn::Count:
            ae.input_type = it;
            ae.sum_type = DataType(TypeId::Int64);
            ae.result_type = DataType(TypeId::Int64);
            break;
          default:
            ae.input_type = it;
            ae.sum_type = it;
            ae.result_type = it;
        }
      }
      n->aggs.push_back(ae);
      if (agg_mode_emits_states(n->agg_mode)) {
        if (ae.fn == AggFn::Avg) {
          n->schema.push_back(Field{ae.name + "[count]", DataType(TypeId::UInt64), true});
          n->schema.push_back(Field{ae.name + "[sum]", ae.sum_type, true});
        } else if (ae.fn == AggFn::Count) {
          n->schema.push_back(Field{ae.name + "[count]", DataType(TypeId::Int64), false});
        } else if (ae.fn == AggFn::Sum) {
          n->schema.push_back(Field{ae.name + "[sum]", ae.sum_type, true});
        } else {
          n->schema.push_back(Field{ae.name + (ae.fn == AggFn::Min ? "[min]" : "[max]"), ae.sum_type, true});
        }
      } else {
        n->schema.push_back(Field{ae.name, ae.result_type, ae.fn != AggFn::Count});
      }
    }
  } else if (op == "HashJoinExec" || op == "SortMergeJoinExec") {
    // SortMergeJoinExec (datafusion.proto:1433, Ballista's default join, extension.rs:683): same matching
    // semantics as the hash join over co-partitioned inputs; its contract adds an output ordered by the
    // join keys (sort_options), which both engines establish by sorting the join result.
    const bool smj = op == "SortMergeJoinExec";
    n->op = PlanNode::HashJoin;
    PlanNode* l = parse_child("left");
    PlanNode* r = parse_child("right");
    n->join_type = parse_join_type(j.get_str("join_type", "Inner"));
    n->partition_mode = smj ? std::string("Partitioned") : j.get_str("mode", "Partitioned");
    n->null_equals_null = j.get_bool("null_equals_null", false);
    const Json& on = j.at("on");
    for (size_t i = 0; i < on.size(); i++) {
      ExprPtr le = parse_expr(on.at(i).at(0), l->schema);
      ExprPtr re = parse_expr(on.at(i).at(1), r->schema);
      if (le->type.pk() != re->type.pk()) throw std::runtime_error("join key types differ: " + le->type.str() + " vs " + re->type.str());
      n->on.emplace_back(le, re);
    }
    Schema both;
    bool lnull = n->join_type == JoinType::Right || n->join_type == JoinType::Full;
    bool rnull = n->join_type == JoinType::Left || n->join_type == JoinType::Full;
    Schema cat = l->schema;
    cat.insert(cat.end(), r->schema.begin(), r->schema.end());
    if (j.has("filter") && j.has("filter_columns")) {
      // the protobuf form (JoinFilter, datafusion.proto:1343-1352): the expression indexes an intermediate schema whose k-th
      // column is (side, index) of an input; rewritten here onto the left ++ right schema the engines evaluate it on
      const Json& fc = j.at("filter_columns");
      Schema inter;
      std::vector<int> target;
      for (size_t k = 0; k < fc.size(); k++) {
        const int side = (int)fc.at(k).at(0).as_int(), idx = (int)fc.at(k).at(1).as_int();
        const Schema& src = side == 0 ? l->schema : r->schema;
        if (side < 0 || side > 1 || idx < 0 || (size_t)idx >= src.size()) throw std::runtime_error("join filter column out of range");
        inter.push_back(src[idx]);
        target.push_back(side == 0 ? idx : (int)l->schema.size() + idx);
      }
      n->join_filter = parse_expr(j.at("filter"), inter);
      remap_columns(n->join_filter, target);
    } else if (j.has("filter")) {
      n->join_filter = parse_expr(j.at("filter"), cat);
    }
    switch (n->join_type) {
      case JoinType::LeftSemi:
      case JoinType::LeftAnti: both = l->schema; break;
      case JoinType::RightSemi:
      case JoinType::RightAnti: both = r->schema; break;
      default:
        for (auto f : l->schema) {
          f.nullable = f.nullable || lnull;
          both.push_back(f);
        }
        for (auto f : r->schema) {
          f.nullable = f.nullable || rnull;
          both.push_back(f);
        }
    }
    if (j.has("projection")) {
      n->has_projection = true;
      const Json& p = j.at("projection");
      for (size_t i = 0; i < p.size(); i++) {
        int idx = (int)p.at(i).as_int();
        if (idx < 0 || (size_t)idx >= both.size()) throw std::runtime_error("join projection out of range");
        n->projection.push_back(idx);
        n->schema.push_back(both[idx]);
      }
    } else {
      n->schema = both;
    }
    if (smj) {
      if (n->has_projection) throw std::runtime_error("SortMergeJoinExec has no projection");
      const bool right_side = n->join_type == JoinType::Right || n->join_type == JoinType::RightSemi || n->join_type == JoinType::RightAnti;
      const int shift = n->join_type == JoinType::Right ? (int)l->schema.size() : 0;
      for (size_t i = 0; i < n->on.size(); i++) {
        SortKey k;
        k.expr = right_side ? shift_cols(n->on[i].second, shift) : n->on[i].first;
        if (j.has("sort_options") && i < j.at("sort_options").size()) {
          const Json& so = j.at("sort_options").at(i);
          k.asc = so.get_bool("asc", true);
          k.nulls_first = so.get_bool("nulls_first", !k.asc);
        }
        n->sort_keys.push_back(k);
      }
    }
  } else if (op == "SortExec" || op == "SortPreservingMergeExec") {
    n->op = op == "SortExec" ? PlanNode::Sort : PlanNode::SortPreservingMerge;
    PlanNode* c = parse_child("input");
    n->schema = c->schema;
    n->sort_keys = parse_sort_keys(j.at("expr"), c->schema);
    n->fetch = j.get_int("fetch", -1);
    n->preserve_partitioning = j.get_bool("preserve_partitioning", false);
  } else if (op == "CoalesceBatchesExec" || op == "CoalescePartitionsExec" || op == "RepartitionExec") {
    n->op = PlanNode::Passthrough;
    PlanNode* c = parse_child("input");
    n->schema = c->schema;
  } else if (op == "GlobalLimitExec" || op == "LocalLimitExec") {
    n->op = PlanNode::Limit;
    PlanNode* c = parse_child("input");
    n->schema = c->schema;
    n->fetch = j.get_int("fetch", -1);
    n->skip = j.get_int("skip", 0);
  } else if (op == "ShuffleWriterExec" || op == "SortShuffleWriterExec") {
    n->op = PlanNode::ShuffleWriter;
    PlanNode* c = parse_child("input");
    n->schema = c->schema;
    n->job_id = j.get_str("job_id", "job");
    n->stage_id = j.get_int("stage_id", 0);
    n->sort_shuffle = (op == "SortShuffleWriterExec") || j.get_bool("sort_shuffle", false);
    if (j.has("partitioning")) {
      const Json& p = j.at("partitioning");
      const Json& hs = p.at("hash");
      for (size_t i = 0; i < hs.size(); i++) n->part_exprs.push_back(parse_expr(hs.at(i), c->schema));
      n->n_out_partitions = p.at("n").as_int();
      if (n->n_out_partitions <= 0) throw std::runtime_error("shuffle writer: partition count must be positive");
    }
  } else {
    throw std::runtime_error("plan IR: unknown operator '" + op + "'");
  }
  return n;
}

}  // namespace b200
