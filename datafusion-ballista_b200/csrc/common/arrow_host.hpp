// Host-side Arrow C Data Interface import/export plumbing (no arithmetic on data).
// A RecordBatch crosses the C-ABI as a struct-typed ArrowArray ("+s") exactly the way
// arrow-rs `arrow::ffi::to_ffi(&StructArray::from(batch).to_data())` and pyarrow
// `RecordBatch._export_to_c` produce it.
#pragma once
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../../include/b200_arrow_abi.h"
#include "plan.hpp"

namespace b200 {

struct ImportedCol {
  std::string name;
  DataType type;
  bool nullable = true;
  const uint8_t* validity = nullptr;  // bitmap (bit i+offset) or null
  const uint8_t* data = nullptr;      // values buffer / offsets buffer (Utf8) / bitmap (Bool)
  const uint8_t* extra = nullptr;     // Utf8 bytes
  int64_t offset = 0;
  int64_t length = 0;
  int64_t null_count = 0;
  bool large_offsets = false;  // LargeUtf8/LargeBinary: int64 offsets
  int ts_unit = 0;
};

inline DataType arrow_format_to_type(const char* f, bool* large, bool* ok) {
  *large = false;
  *ok = true;
  std::string s(f);
  if (s == "n") return DataType(TypeId::Null);
  if (s == "b") return DataType(TypeId::Bool);
  if (s == "c") return DataType(TypeId::Int8);
  if (s == "C") return DataType(TypeId::UInt8);
  if (s == "s") return DataType(TypeId::Int16);
  if (s == "S") return DataType(TypeId::UInt16);
  if (s == "i") return DataType(TypeId::Int32);
  if (s == "I") return DataType(TypeId::UInt32);
  if (s == "l") return DataType(TypeId::Int64);
  if (s == "L") return DataType(TypeId::UInt64);
  if (s == "f") return DataType(TypeId::Float32);
  if (s == "g") return DataType(TypeId::Float64);
  if (s == "tdD") return DataType(TypeId::Date32);
  if (s.rfind("ts", 0) == 0) return DataType(TypeId::Timestamp);
  if (s == "u" || s == "z") return DataType(TypeId::Utf8);
  if (s == "U" || s == "Z") {
    *large = true;
    return DataType(TypeId::Utf8);
  }
  if (s.rfind("d:", 0) == 0) {
    int p = 0, sc = 0, bits = 128;
    int n = sscanf(f, "d:%d,%d,%d", &p, &sc, &bits);
    if (n >= 2 && bits == 128 && p >= 1 && p <= 38) return DataType::decimal(p, sc);
  }
  *ok = false;
  return DataType();
}

inline std::string type_to_arrow_format(const DataType& t) {
  switch (t.id) {
    case TypeId::Null: return "n";
    case TypeId::Bool: return "b";
    case TypeId::Int8: return "c";
    case TypeId::UInt8: return "C";
    case TypeId::Int16: return "s";
    case TypeId::UInt16: return "S";
    case TypeId::Int32: return "i";
    case TypeId::UInt32: return "I";
    case TypeId::Int64: return "l";
    case TypeId::UInt64: return "L";
    case TypeId::Float32: return "f";
    case TypeId::Float64: return "g";
    case TypeId::Date32: return "tdD";
    case TypeId::Timestamp: return "tsn:";
    case TypeId::Utf8: return "u";
    case TypeId::Decimal128:
      return "d:" + std::to_string((int)t.precision) + "," + std::to_string((int)t.scale);
  }
  return "n";
}

// Import a struct array (record batch). Throws std::runtime_error on unsupported layouts.
inline std::vector<ImportedCol> import_record_batch(const ArrowArray* arr, const ArrowSchema* sch, int64_t* n_rows) {
  if (!arr || !sch || !sch->format) throw std::runtime_error("import: null ArrowArray/ArrowSchema");
  if (std::string(sch->format) != "+s") throw std::runtime_error("import: expected a struct array (record batch)");
  if (arr->n_children != sch->n_children) throw std::runtime_error("import: children count mismatch");
  if (arr->offset != 0) throw std::runtime_error("import: sliced struct arrays are not supported");
  *n_rows = arr->length;
  std::vector<ImportedCol> out;
  for (int64_t c = 0; c < arr->n_children; c++) {
    const ArrowArray* a = arr->children[c];
    const ArrowSchema* s = sch->children[c];
    ImportedCol ic;
    ic.name = s->name ? s->name : "";
    bool ok = false;
    ic.type = arrow_format_to_type(s->format, &ic.large_offsets, &ok);
    if (!ok) throw std::runtime_error(std::string("import: unsupported Arrow format '") + s->format + "' for column " + ic.name);
    if (a->dictionary) throw std::runtime_error("import: dictionary arrays are not supported");
    ic.nullable = (s->flags & ARROW_FLAG_NULLABLE) != 0;
    ic.offset = a->offset;
    ic.length = a->length;
    if (a->length != arr->length) throw std::runtime_error("import: child length mismatch");
    ic.null_count = a->null_count;
    if (ic.type.id == TypeId::Null) {
      ic.null_count = a->length;
    } else {
      if (a->n_buffers < 2) throw std::runtime_error("import: missing buffers");
      ic.validity = (const uint8_t*)a->buffers[0];
      ic.data = (const uint8_t*)a->buffers[1];
      if (ic.type.id == TypeId::Utf8) {
        if (a->n_buffers < 3) throw std::runtime_error("import: utf8 needs 3 buffers");
        ic.extra = (const uint8_t*)a->buffers[2];
      }
      if (ic.null_count < 0) {
        // unknown: count
        int64_t nc = 0;
        if (ic.validity)
          for (int64_t i = 0; i < a->length; i++) {
            int64_t b = i + a->offset;
            nc += !((ic.validity[b >> 3] >> (b & 7)) & 1);
          }
        ic.null_count = nc;
      }
      if (!ic.validity) ic.null_count = 0;
    }
    out.push_back(ic);
  }
  return out;
}

// ---------------------------------------------------------------------------------------------
// Export
// ---------------------------------------------------------------------------------------------
struct HostCol {
  std::string name;
  DataType type;
  bool nullable = true;
  int64_t n = 0;
  int64_t null_count = 0;
  std::vector<uint8_t> validity;  // bitmap; empty => all valid
  std::vector<uint8_t> data;      // values (or int32 offsets for Utf8, bitmap for Bool)
  std::vector<uint8_t> extra;     // Utf8 bytes
};

namespace detail {
struct ExportPriv {
  std::vector<HostCol> cols;
  std::vector<std::string> formats;
  std::vector<ArrowArray> child_arrays;
  std::vector<ArrowArray*> child_array_ptrs;
  std::vector<std::vector<const void*>> child_buffers;
  std::vector<ArrowSchema> child_schemas;
  std::vector<ArrowSchema*> child_schema_ptrs;
  const void* top_buffers[1] = {nullptr};
};
inline void release_child_array(ArrowArray* a) { a->release = nullptr; }
inline void release_child_schema(ArrowSchema* s) { s->release = nullptr; }
inline void release_top_array(ArrowArray* a) {
  delete (ExportPriv*)a->private_data;
  a->release = nullptr;
}
inline void release_top_schema(ArrowSchema* s) {
  delete (ExportPriv*)s->private_data;
  s->release = nullptr;
}
}  // namespace detail

// Moves `cols` into heap storage owned by the exported ArrowArray; the schema gets its own copy of
// names/formats.  Consumer calls release on both (standard ownership rule).
inline void export_record_batch(std::vector<HostCol>&& cols, int64_t n_rows, ArrowArray* out, ArrowSchema* out_schema) {
  using namespace detail;
  size_t nc = cols.size();
  if (out_schema) {
    auto* sp = new ExportPriv();
    sp->formats.resize(nc);
    sp->child_schemas.resize(nc);
    sp->child_schema_ptrs.resize(nc);
    sp->cols.resize(nc);
    for (size_t i = 0; i < nc; i++) {
      sp->cols[i].name = cols[i].name;
      sp->formats[i] = type_to_arrow_format(cols[i].type);
      ArrowSchema& s = sp->child_schemas[i];
      memset(&s, 0, sizeof s);
      s.format = sp->formats[i].c_str();
      s.name = sp->cols[i].name.c_str();
      s.flags = cols[i].nullable ? ARROW_FLAG_NULLABLE : 0;
      s.release = release_child_schema;
      sp->child_schema_ptrs[i] = &s;
    }
    memset(out_schema, 0, sizeof *out_schema);
    out_schema->format = "+s";
    out_schema->name = "";
    out_schema->n_children = (int64_t)nc;
    out_schema->children = sp->child_schema_ptrs.data();
    out_schema->release = release_top_schema;
    out_schema->private_data = sp;
  }
  if (out) {
    auto* ap = new ExportPriv();
    ap->cols = std::move(cols);
    ap->child_arrays.resize(nc);
    ap->child_array_ptrs.resize(nc);
    ap->child_buffers.resize(nc);
    for (size_t i = 0; i < nc; i++) {
      HostCol& c = ap->cols[i];
      ArrowArray& a = ap->child_arrays[i];
      memset(&a, 0, sizeof a);
      a.length = c.n;
      a.null_count = c.null_count;
      auto& bufs = ap->child_buffers[i];
      static const uint8_t kEmpty[16] = {0};
      if (c.type.id == TypeId::Null) {
        a.n_buffers = 0;
      } else {
        bufs.push_back(c.validity.empty() ? nullptr : c.validity.data());
        bufs.push_back(c.data.empty() ? (const void*)kEmpty : c.data.data());
        if (c.type.id == TypeId::Utf8) bufs.push_back(c.extra.empty() ? (const void*)kEmpty : c.extra.data());
        a.n_buffers = (int64_t)bufs.size();
      }
      a.buffers = bufs.data();
      a.release = release_child_array;
      ap->child_array_ptrs[i] = &a;
    }
    memset(out, 0, sizeof *out);
    out->length = n_rows;
    out->null_count = 0;
    out->n_buffers = 1;
    out->buffers = ap->top_buffers;
    out->n_children = (int64_t)nc;
    out->children = ap->child_array_ptrs.data();
    out->release = release_top_array;
    out->private_data = ap;
  }
}

}  // namespace b200
