// Arrow IPC streams with LZ4_FRAME body compression: the byte format of the reference's shuffle files.
//
// Reference writers / readers this is compatible with:
//   hash shuffle      ShuffleWriterExec: one `StreamWriter` (IpcWriteOptions + LZ4_FRAME) per output partition,
//                     work_dir/job/stage/{out_part}/data-{in_part}.arrow        (shuffle_writer.rs:317-328, mod.rs:66-99)
//   sort shuffle      one data file = [schema-only stream][per partition: concatenated IPC streams] + `data.arrow.index`
//                     of (P+1) little-endian i64 offsets                         (sort_shuffle/writer.rs:419-513, index.rs:18-33)
//   readers           arrow-rs `StreamReader` (shuffle_reader.rs:760-768, client.rs:515-517, sort_shuffle/reader.rs:51-84)
// Everything here is restated from the published formats, nothing is linked: Arrow columnar IPC (encapsulated message
// framing, Message.fbs / Schema.fbs field ids), FlatBuffers binary layout (tables, vtables, vectors, strings), the LZ4
// frame and block formats and xxHash32 (frame header checksum).  Host only, no CUDA: the device side exports / imports
// HostCol batches (arrow_host.hpp).  Pinned by tests/test_ipc_format.py against pyarrow's reader and writer.
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../common/arrow_host.hpp"

namespace b200 {
namespace ipc {

// ------------------------------------------------------------------------------------------------
// xxHash32 + LZ4 block / frame
// ------------------------------------------------------------------------------------------------
inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
inline uint32_t rd32(const uint8_t* p) {
  uint32_t v;
  memcpy(&v, p, 4);
  return v;
}
inline uint32_t xxh32(const uint8_t* p, size_t len, uint32_t seed) {
  const uint32_t P1 = 2654435761u, P2 = 2246822519u, P3 = 3266489917u, P4 = 668265263u, P5 = 374761393u;
  const uint8_t* end = p + len;
  uint32_t h;
  if (len >= 16) {
    uint32_t v1 = seed + P1 + P2, v2 = seed + P2, v3 = seed, v4 = seed - P1;
    const uint8_t* limit = end - 16;
    do {
      v1 = rotl32(v1 + rd32(p) * P2, 13) * P1;
      v2 = rotl32(v2 + rd32(p + 4) * P2, 13) * P1;
      v3 = rotl32(v3 + rd32(p + 8) * P2, 13) * P1;
      v4 = rotl32(v4 + rd32(p + 12) * P2, 13) * P1;
      p += 16;
    } while (p <= limit);
    h = rotl32(v1, 1) + rotl32(v2, 7) + rotl32(v3, 12) + rotl32(v4, 18);
  } else {
    h = seed + P5;
  }
  h += (uint32_t)len;
  while (p + 4 <= end) {
    h = rotl32(h + rd32(p) * P3, 17) * P4;
    p += 4;
  }
  while (p < end) {
    h = rotl32(h + (*p) * P5, 11) * P1;
    p++;
  }
  h ^= h >> 15;
  h *= P2;
  h ^= h >> 13;
  h *= P3;
  h ^= h >> 16;
  return h;
}

// LZ4 block decode into dst[0, cap); `window` = start of the output the block may reference (the frame's output so far for
// linked blocks, else dst itself); returns bytes produced or throws
inline size_t lz4_block_decode(const uint8_t* src, size_t n, uint8_t* dst, size_t cap, const uint8_t* window = nullptr) {
  if (!window) window = dst;
  const uint8_t* ip = src;
  const uint8_t* iend = src + n;
  uint8_t* op = dst;
  uint8_t* oend = dst + cap;
  while (ip < iend) {
    const uint8_t token = *ip++;
    size_t lit = token >> 4;
    if (lit == 15) {
      uint8_t b;
      do {
        if (ip >= iend) throw std::runtime_error("lz4: truncated literal length");
        b = *ip++;
        lit += b;
      } while (b == 255);
    }
    if ((size_t)(iend - ip) < lit || (size_t)(oend - op) < lit) throw std::runtime_error("lz4: literal overrun");
    memcpy(op, ip, lit);
    ip += lit;
    op += lit;
    if (ip >= iend) break;  // the last sequence carries literals only
    if (iend - ip < 2) throw std::runtime_error("lz4: truncated offset");
    const size_t off = (size_t)ip[0] | ((size_t)ip[1] << 8);
    ip += 2;
    if (off == 0 || off > (size_t)(op - window)) throw std::runtime_error("lz4: bad match offset");
    size_t ml = token & 15;
    if (ml == 15) {
      uint8_t b;
      do {
        if (ip >= iend) throw std::runtime_error("lz4: truncated match length");
        b = *ip++;
        ml += b;
      } while (b == 255);
    }
    ml += 4;
    if ((size_t)(oend - op) < ml) throw std::runtime_error("lz4: match overrun");
    const uint8_t* m = op - off;
    for (size_t k = 0; k < ml; k++) op[k] = m[k];  // overlapping copies are part of the format
    op += ml;
  }
  return (size_t)(op - dst);
}

// greedy LZ4 block encoder (hash of 4 bytes -> last position); output is a valid block for any input
inline void lz4_block_encode(const uint8_t* src, size_t n, std::vector<uint8_t>& out) {
  const size_t MINMATCH = 4, MFLIMIT = 12, LASTLITERALS = 5;
  std::vector<int32_t> table(1 << 16, -1);
  size_t anchor = 0, i = 0;
  auto emit = [&](size_t lit_start, size_t lit_len, size_t match_len, size_t offset) {
    const size_t ml = match_len ? match_len - MINMATCH : 0;
    uint8_t token = (uint8_t)((lit_len >= 15 ? 15 : lit_len) << 4) | (uint8_t)(match_len ? (ml >= 15 ? 15 : ml) : 0);
    out.push_back(token);
    if (lit_len >= 15) {
      size_t r = lit_len - 15;
      while (r >= 255) {
        out.push_back(255);
        r -= 255;
      }
      out.push_back((uint8_t)r);
    }
    out.insert(out.end(), src + lit_start, src + lit_start + lit_len);
    if (match_len) {
      out.push_back((uint8_t)(offset & 0xFF));
      out.push_back((uint8_t)(offset >> 8));
      if (ml >= 15) {
        size_t r = ml - 15;
        while (r >= 255) {
          out.push_back(255);
          r -= 255;
        }
        out.push_back((uint8_t)r);
      }
    }
  };
  if (n >= MFLIMIT + 1) {
    const size_t mlimit = n - MFLIMIT;
    while (i < mlimit) {
      const uint32_t h = (rd32(src + i) * 2654435761u) >> 16;
      const int32_t cand = table[h];
      table[h] = (int32_t)i;
      if (cand >= 0 && i - (size_t)cand <= 65535 && rd32(src + cand) == rd32(src + i)) {
        size_t ml = MINMATCH;
        const size_t max_ml = n - LASTLITERALS - i;
        while (ml < max_ml && src[(size_t)cand + ml] == src[i + ml]) ml++;
        emit(anchor, i - anchor, ml, i - (size_t)cand);
        i += ml;
        anchor = i;
      } else {
        i++;
      }
    }
  }
  emit(anchor, n - anchor, 0, 0);
}

static const uint32_t LZ4F_MAGIC = 0x184D2204u;
static const size_t LZ4F_BLOCK = 4u << 20;  // BD code 7

inline void lz4_frame_encode(const uint8_t* src, size_t n, std::vector<uint8_t>& out) {
  uint8_t hdr[7];
  memcpy(hdr, &LZ4F_MAGIC, 4);
  hdr[4] = 0x60;  // version 01, block independence, no checksums, no content size
  hdr[5] = 0x70;  // 4 MiB blocks
  hdr[6] = (uint8_t)((xxh32(hdr + 4, 2, 0) >> 8) & 0xFF);
  out.insert(out.end(), hdr, hdr + 7);
  std::vector<uint8_t> blk;
  for (size_t pos = 0; pos < n; pos += LZ4F_BLOCK) {
    const size_t len = n - pos < LZ4F_BLOCK ? n - pos : LZ4F_BLOCK;
    blk.clear();
    lz4_block_encode(src + pos, len, blk);
    uint32_t word;
    const uint8_t* payload;
    size_t plen;
    if (blk.size() >= len) {  // incompressible: stored block
      word = (uint32_t)len | 0x80000000u;
      payload = src + pos;
      plen = len;
    } else {
      word = (uint32_t)blk.size();
      payload = blk.data();
      plen = blk.size();
    }
    const uint8_t* w = (const uint8_t*)&word;
    out.insert(out.end(), w, w + 4);
    out.insert(out.end(), payload, payload + plen);
  }
  const uint32_t endmark = 0;
  const uint8_t* w = (const uint8_t*)&endmark;
  out.insert(out.end(), w, w + 4);
}

inline void lz4_frame_decode(const uint8_t* src, size_t n, uint8_t* dst, size_t dst_len) {
  if (n < 7 || rd32(src) != LZ4F_MAGIC) throw std::runtime_error("lz4 frame: bad magic");
  const uint8_t flg = src[4];
  if ((flg >> 6) != 1) throw std::runtime_error("lz4 frame: unsupported version");
  const bool block_checksum = (flg >> 4) & 1, content_size = (flg >> 3) & 1, dict_id = flg & 1;
  size_t p = 6 + (content_size ? 8 : 0) + (dict_id ? 4 : 0) + 1;  // ... + header checksum byte
  size_t produced = 0;
  for (;;) {
    if (p + 4 > n) throw std::runtime_error("lz4 frame: truncated block header");
    const uint32_t word = rd32(src + p);
    p += 4;
    if (word == 0) break;
    const bool stored = (word & 0x80000000u) != 0;
    const size_t len = word & 0x7FFFFFFFu;
    if (p + len > n) throw std::runtime_error("lz4 frame: truncated block");
    if (stored) {
      if (produced + len > dst_len) throw std::runtime_error("lz4 frame: output overrun");
      memcpy(dst + produced, src + p, len);
      produced += len;
    } else {
      produced += lz4_block_decode(src + p, len, dst + produced, dst_len - produced, dst);  // linked blocks reach back into the frame
    }
    p += len + (block_checksum ? 4 : 0);
  }
  if (produced != dst_len) throw std::runtime_error("lz4 frame: decoded size differs from the recorded buffer length");
}

// ------------------------------------------------------------------------------------------------
// FlatBuffers: a small back-to-front builder and a table reader
// ------------------------------------------------------------------------------------------------
struct FbBuilder {
  std::vector<uint8_t> buf;
  size_t head;
  size_t max_align = 1;
  std::vector<std::pair<int, uint32_t>> fields;  // current table: (field id, offset-from-end of the value)
  uint32_t table_end = 0;
  FbBuilder() : buf(1024), head(1024) {}
  uint32_t size() const { return (uint32_t)(buf.size() - head); }
  void make_room(size_t need) {
    if (head >= need) return;
    const size_t old = buf.size(), used = old - head;
    size_t cap = old * 2;
    while (cap - used < need) cap *= 2;
    std::vector<uint8_t> nb(cap);
    memcpy(nb.data() + cap - used, buf.data() + head, used);
    buf.swap(nb);
    head = cap - used;
  }
  void pad(size_t n) {
    make_room(n);
    head -= n;
    memset(buf.data() + head, 0, n);
  }
  // make (size() + additional) a multiple of a
  void align(size_t a, size_t additional = 0) {
    if (a > max_align) max_align = a;
    const size_t rem = (size() + additional) % a;
    if (rem) pad(a - rem);
  }
  void push(const void* p, size_t n) {
    make_room(n);
    head -= n;
    memcpy(buf.data() + head, p, n);
  }
  template <class T>
  uint32_t push_scalar(T v) {
    align(sizeof(T));
    push(&v, sizeof(T));
    return size();
  }
  uint32_t create_string(const std::string& s) {
    align(4, s.size() + 1);
    pad(1);
    push(s.data(), s.size());
    uint32_t len = (uint32_t)s.size();
    push(&len, 4);
    return size();
  }
  // vector of structs / scalars given as raw bytes
  uint32_t create_vector_raw(const void* data, size_t elem_size, size_t count, size_t elem_align) {
    align(4, elem_size * count);
    align(elem_align, elem_size * count);
    push(data, elem_size * count);
    uint32_t len = (uint32_t)count;
    align(4);
    push(&len, 4);
    return size();
  }
  uint32_t create_vector_of_offsets(const std::vector<uint32_t>& offs) {
    align(4, 4 * offs.size());
    for (size_t i = offs.size(); i-- > 0;) {
      const uint32_t rel = size() + 4 - offs[i];
      push(&rel, 4);
    }
    uint32_t len = (uint32_t)offs.size();
    push(&len, 4);
    return size();
  }
  void start_table() {
    fields.clear();
    table_end = size();
  }
  template <class T>
  void add_scalar(int id, T v) {
    fields.push_back({id, push_scalar<T>(v)});
  }
  void add_offset(int id, uint32_t target) {
    align(4);
    const uint32_t rel = size() + 4 - target;
    push(&rel, 4);
    fields.push_back({id, size()});
  }
  uint32_t end_table() {
    align(4);
    int32_t placeholder = 0;
    push(&placeholder, 4);
    const uint32_t table_off = size();
    int max_id = -1;
    for (auto& f : fields) max_id = f.first > max_id ? f.first : max_id;
    std::vector<uint16_t> vt((size_t)(max_id + 1) + 2, 0);
    vt[0] = (uint16_t)(vt.size() * 2);
    vt[1] = (uint16_t)(table_off - table_end);
    for (auto& f : fields) vt[(size_t)f.first + 2] = (uint16_t)(table_off - f.second);
    align(2, 0);
    // the vtable goes right before the table
    push(vt.data(), vt.size() * 2);
    const uint32_t vt_off = size();
    const int32_t soffset = (int32_t)(vt_off - table_off);
    memcpy(buf.data() + buf.size() - table_off, &soffset, 4);
    return table_off;
  }
  // finished buffer (root offset first), padded so that its length is a multiple of 8
  std::vector<uint8_t> finish(uint32_t root) {
    align(max_align < 8 ? 8 : max_align, 4);
    const uint32_t rel = size() + 4 - root;
    push(&rel, 4);
    return std::vector<uint8_t>(buf.begin() + (long)head, buf.end());
  }
};

struct FbTable {
  const uint8_t* buf = nullptr;
  size_t len = 0, pos = 0;
  bool ok() const { return buf != nullptr; }
  static FbTable root(const uint8_t* b, size_t n) {
    if (n < 8) throw std::runtime_error("flatbuffer: too short");
    FbTable t;
    t.buf = b;
    t.len = n;
    t.pos = rd32(b);
    if (t.pos + 4 > n) throw std::runtime_error("flatbuffer: bad root");
    return t;
  }
  size_t field_pos(int id) const {
    int32_t so;
    memcpy(&so, buf + pos, 4);
    const size_t vt = (size_t)((int64_t)pos - so);
    if (vt + 4 > len) throw std::runtime_error("flatbuffer: bad vtable");
    uint16_t vsize;
    memcpy(&vsize, buf + vt, 2);
    const size_t slot = 4 + 2 * (size_t)id;
    if (slot + 2 > vsize) return 0;
    uint16_t off;
    memcpy(&off, buf + vt + slot, 2);
    return off ? pos + off : 0;
  }
  template <class T>
  T scalar(int id, T dflt) const {
    const size_t p = field_pos(id);
    if (!p) return dflt;
    T v;
    memcpy(&v, buf + p, sizeof(T));
    return v;
  }
  size_t indirect(int id) const {  // position a uoffset field points at
    const size_t p = field_pos(id);
    if (!p) return 0;
    return p + rd32(buf + p);
  }
  FbTable table(int id) const {
    FbTable t;
    const size_t p = indirect(id);
    if (!p) return t;
    t.buf = buf;
    t.len = len;
    t.pos = p;
    return t;
  }
  std::string str(int id) const {
    const size_t p = indirect(id);
    if (!p) return "";
    const uint32_t n = rd32(buf + p);
    return std::string((const char*)buf + p + 4, n);
  }
  // vector: returns element count, *first = position of element 0
  uint32_t vec(int id, size_t* first) const {
    const size_t p = indirect(id);
    if (!p) {
      *first = 0;
      return 0;
    }
    *first = p + 4;
    return rd32(buf + p);
  }
  FbTable vec_table(size_t first, uint32_t i) const {
    FbTable t;
    const size_t e = first + 4 * (size_t)i;
    t.buf = buf;
    t.len = len;
    t.pos = e + rd32(buf + e);
    return t;
  }
};

// ------------------------------------------------------------------------------------------------
// Arrow IPC stream writer / reader over HostCol batches
// ------------------------------------------------------------------------------------------------
enum { MSG_SCHEMA = 1, MSG_RECORD_BATCH = 3 };
enum { TY_NULL = 1, TY_INT = 2, TY_FLOAT = 3, TY_UTF8 = 5, TY_BOOL = 6, TY_DECIMAL = 7, TY_DATE = 8, TY_TIMESTAMP = 10, TY_LARGE_UTF8 = 20 };

inline uint32_t fb_type(FbBuilder& b, const DataType& t, uint8_t* tag) {
  auto int_type = [&](int bits, bool sgn) {
    b.start_table();
    b.add_scalar<int32_t>(0, bits);
    b.add_scalar<uint8_t>(1, sgn ? 1 : 0);
    *tag = TY_INT;
    return b.end_table();
  };
  switch (t.id) {
    case TypeId::Int8: return int_type(8, true);
    case TypeId::Int16: return int_type(16, true);
    case TypeId::Int32: return int_type(32, true);
    case TypeId::Int64: return int_type(64, true);
    case TypeId::UInt8: return int_type(8, false);
    case TypeId::UInt16: return int_type(16, false);
    case TypeId::UInt32: return int_type(32, false);
    case TypeId::UInt64: return int_type(64, false);
    case TypeId::Float32:
    case TypeId::Float64:
      b.start_table();
      b.add_scalar<int16_t>(0, t.id == TypeId::Float32 ? 1 : 2);
      *tag = TY_FLOAT;
      return b.end_table();
    case TypeId::Utf8:
      b.start_table();
      *tag = TY_UTF8;
      return b.end_table();
    case TypeId::Bool:
      b.start_table();
      *tag = TY_BOOL;
      return b.end_table();
    case TypeId::Decimal128:
      b.start_table();
      b.add_scalar<int32_t>(0, t.precision);
      b.add_scalar<int32_t>(1, t.scale);
      b.add_scalar<int32_t>(2, 128);
      *tag = TY_DECIMAL;
      return b.end_table();
    case TypeId::Date32:
      b.start_table();
      b.add_scalar<int16_t>(0, 0);  // DateUnit::DAY
      *tag = TY_DATE;
      return b.end_table();
    case TypeId::Timestamp:
      b.start_table();
      b.add_scalar<int16_t>(0, 3);  // NANOSECOND
      *tag = TY_TIMESTAMP;
      return b.end_table();
    default:
      b.start_table();
      *tag = TY_NULL;
      return b.end_table();
  }
}

inline void put_message(std::vector<uint8_t>& out, const std::vector<uint8_t>& meta, const std::vector<uint8_t>& body) {
  const uint32_t cont = 0xFFFFFFFFu;
  const int32_t msize = (int32_t)meta.size();  // FbBuilder::finish pads to 8
  out.insert(out.end(), (const uint8_t*)&cont, (const uint8_t*)&cont + 4);
  out.insert(out.end(), (const uint8_t*)&msize, (const uint8_t*)&msize + 4);
  out.insert(out.end(), meta.begin(), meta.end());
  out.insert(out.end(), body.begin(), body.end());
}

inline void write_schema(std::vector<uint8_t>& out, const std::vector<HostCol>& cols) {
  FbBuilder b;
  std::vector<uint32_t> fields;
  for (auto& c : cols) {
    uint8_t tag = 0;
    const uint32_t ty = fb_type(b, c.type, &tag);
    const uint32_t name = b.create_string(c.name);
    const uint32_t children = b.create_vector_of_offsets({});
    b.start_table();
    b.add_offset(0, name);
    b.add_scalar<uint8_t>(1, c.nullable ? 1 : 0);
    b.add_scalar<uint8_t>(2, tag);
    b.add_offset(3, ty);
    b.add_offset(5, children);
    fields.push_back(b.end_table());
  }
  const uint32_t fv = b.create_vector_of_offsets(fields);
  b.start_table();
  b.add_scalar<int16_t>(0, 0);  // little endian
  b.add_offset(1, fv);
  const uint32_t schema = b.end_table();
  b.start_table();
  b.add_scalar<int16_t>(0, 4);  // MetadataVersion::V5
  b.add_scalar<uint8_t>(1, MSG_SCHEMA);
  b.add_offset(2, schema);
  b.add_scalar<int64_t>(3, 0);
  const uint32_t msg = b.end_table();
  put_message(out, b.finish(msg), {});
}

inline void write_eos(std::vector<uint8_t>& out) {
  const uint32_t w[2] = {0xFFFFFFFFu, 0u};
  out.insert(out.end(), (const uint8_t*)w, (const uint8_t*)w + 8);
}

// one RecordBatch message; compress: LZ4_FRAME per buffer (a buffer that does not shrink is stored with length -1)
inline void write_batch(std::vector<uint8_t>& out, const std::vector<HostCol>& cols, int64_t n_rows, bool compress) {
  struct Node { int64_t length, null_count; };
  struct Buf { int64_t offset, length; };
  std::vector<Node> nodes;
  std::vector<Buf> bufs;
  std::vector<uint8_t> body;
  auto add_buffer = [&](const uint8_t* p, size_t n) {
    Buf bf;
    bf.offset = (int64_t)body.size();
    if (n == 0) {
      bf.length = 0;
    } else if (!compress) {
      body.insert(body.end(), p, p + n);
      bf.length = (int64_t)n;
    } else {
      std::vector<uint8_t> frame;
      lz4_frame_encode(p, n, frame);
      int64_t ulen = (int64_t)n;
      if (frame.size() >= n) {  // not worth it: stored, marked with -1
        ulen = -1;
        body.insert(body.end(), (const uint8_t*)&ulen, (const uint8_t*)&ulen + 8);
        body.insert(body.end(), p, p + n);
        bf.length = 8 + (int64_t)n;
      } else {
        body.insert(body.end(), (const uint8_t*)&ulen, (const uint8_t*)&ulen + 8);
        body.insert(body.end(), frame.begin(), frame.end());
        bf.length = 8 + (int64_t)frame.size();
      }
    }
    while (body.size() % 8) body.push_back(0);
    bufs.push_back(bf);
  };
  for (auto& c : cols) {
    nodes.push_back(Node{n_rows, c.null_count});
    if (c.type.id == TypeId::Null) continue;
    if (c.null_count > 0 && !c.validity.empty()) add_buffer(c.validity.data(), c.validity.size());
    else add_buffer(nullptr, 0);
    add_buffer(c.data.data(), c.data.size());
    if (c.type.id == TypeId::Utf8) add_buffer(c.extra.data(), c.extra.size());
  }
  FbBuilder b;
  uint32_t comp = 0;
  if (compress) {
    b.start_table();
    b.add_scalar<int8_t>(0, 0);  // CompressionType::LZ4_FRAME
    b.add_scalar<int8_t>(1, 0);  // BodyCompressionMethod::BUFFER
    comp = b.end_table();
  }
  const uint32_t bv = b.create_vector_raw(bufs.data(), sizeof(Buf), bufs.size(), 8);
  const uint32_t nv = b.create_vector_raw(nodes.data(), sizeof(Node), nodes.size(), 8);
  b.start_table();
  b.add_scalar<int64_t>(0, n_rows);
  b.add_offset(1, nv);
  b.add_offset(2, bv);
  if (compress) b.add_offset(3, comp);
  const uint32_t rb = b.end_table();
  b.start_table();
  b.add_scalar<int16_t>(0, 4);
  b.add_scalar<uint8_t>(1, MSG_RECORD_BATCH);
  b.add_offset(2, rb);
  b.add_scalar<int64_t>(3, (int64_t)body.size());
  const uint32_t msg = b.end_table();
  put_message(out, b.finish(msg), body);
}

// A complete stream: schema, the batch cut into pieces of at most `max_rows` rows (0: one message), end-of-stream
inline void write_stream(std::vector<uint8_t>& out, const std::vector<HostCol>& cols, int64_t n_rows, bool compress, int64_t max_rows = 0) {
  write_schema(out, cols);
  if (max_rows <= 0 || n_rows <= max_rows) {
    if (n_rows > 0) write_batch(out, cols, n_rows, compress);
  } else {
    for (int64_t r0 = 0; r0 < n_rows; r0 += max_rows) {
      const int64_t r1 = r0 + max_rows < n_rows ? r0 + max_rows : n_rows;
      std::vector<HostCol> part;
      for (auto& c : cols) {
        HostCol p;
        p.name = c.name;
        p.type = c.type;
        p.nullable = c.nullable;
        p.n = r1 - r0;
        auto bit = [](const std::vector<uint8_t>& bm, int64_t i) { return (bm[(size_t)(i >> 3)] >> (i & 7)) & 1; };
        if (c.null_count > 0 && !c.validity.empty()) {
          p.validity.assign((size_t)((p.n + 7) / 8), 0);
          for (int64_t i = 0; i < p.n; i++)
            if (bit(c.validity, r0 + i)) p.validity[(size_t)(i >> 3)] |= (uint8_t)(1u << (i & 7));
            else p.null_count++;
          if (p.null_count == 0) p.validity.clear();
        }
        if (c.type.id == TypeId::Bool) {
          p.data.assign((size_t)((p.n + 7) / 8), 0);
          for (int64_t i = 0; i < p.n; i++)
            if (bit(c.data, r0 + i)) p.data[(size_t)(i >> 3)] |= (uint8_t)(1u << (i & 7));
        } else if (c.type.id == TypeId::Utf8) {
          const int32_t* off = (const int32_t*)c.data.data();
          p.data.resize((size_t)(p.n + 1) * 4);
          int32_t* po = (int32_t*)p.data.data();
          for (int64_t i = 0; i <= p.n; i++) po[i] = off[r0 + i] - off[r0];
          p.extra.assign(c.extra.begin() + off[r0], c.extra.begin() + off[r1]);
        } else if (c.type.id != TypeId::Null) {
          const size_t w = c.data.size() / (size_t)(c.n ? c.n : 1);
          p.data.assign(c.data.begin() + (long)((size_t)r0 * w), c.data.begin() + (long)((size_t)r1 * w));
        }
        part.push_back(std::move(p));
      }
      write_batch(out, part, r1 - r0, compress);
    }
  }
  write_eos(out);
}

inline DataType read_type(const FbTable& field) {
  const uint8_t tag = field.scalar<uint8_t>(2, 0);
  const FbTable ty = field.table(3);
  switch (tag) {
    case TY_INT: {
      const int bits = ty.ok() ? ty.scalar<int32_t>(0, 0) : 0;
      const bool sgn = ty.ok() && ty.scalar<uint8_t>(1, 0);
      switch (bits) {
        case 8: return DataType(sgn ? TypeId::Int8 : TypeId::UInt8);
        case 16: return DataType(sgn ? TypeId::Int16 : TypeId::UInt16);
        case 32: return DataType(sgn ? TypeId::Int32 : TypeId::UInt32);
        case 64: return DataType(sgn ? TypeId::Int64 : TypeId::UInt64);
      }
      break;
    }
    case TY_FLOAT: {
      const int p = ty.ok() ? ty.scalar<int16_t>(0, 0) : 0;
      if (p == 1) return DataType(TypeId::Float32);
      if (p == 2) return DataType(TypeId::Float64);
      break;
    }
    case TY_UTF8: return DataType(TypeId::Utf8);
    case TY_BOOL: return DataType(TypeId::Bool);
    case TY_DECIMAL:
      if (ty.ok() && ty.scalar<int32_t>(2, 128) == 128) return DataType::decimal(ty.scalar<int32_t>(0, 0), ty.scalar<int32_t>(1, 0));
      break;
    case TY_DATE:
      if (ty.ok() && ty.scalar<int16_t>(0, 1) == 0) return DataType(TypeId::Date32);
      break;
    case TY_TIMESTAMP: return DataType(TypeId::Timestamp);
    case TY_NULL: return DataType(TypeId::Null);
  }
  throw std::runtime_error("arrow ipc: column '" + field.str(0) + "' has a type the engine does not carry (type tag " + std::to_string(tag) + ")");
}

// Reads every message of one or several back-to-back IPC streams in [p, p + n): the schema comes from the first Schema
// message (later ones must agree in column count), all record batches are appended row-wise into `cols`.
inline int64_t read_streams(const uint8_t* p, size_t n, std::vector<HostCol>& cols) {
  size_t pos = 0;
  bool have_schema = !cols.empty();
  int64_t total_rows = cols.empty() ? 0 : cols[0].n;
  while (pos + 8 <= n) {
    uint32_t w0 = rd32(p + pos);
    size_t msize;
    if (w0 == 0xFFFFFFFFu) {
      msize = rd32(p + pos + 4);
      pos += 8;
    } else {  // pre-0.15 framing without the continuation marker
      msize = w0;
      pos += 4;
    }
    if (msize == 0) continue;  // end-of-stream marker: another stream may follow (sort-shuffle partitions)
    if (pos + msize > n) throw std::runtime_error("arrow ipc: truncated message metadata");
    const FbTable msg = FbTable::root(p + pos, msize);
    const uint8_t htype = msg.scalar<uint8_t>(1, 0);
    const int64_t body_len = msg.scalar<int64_t>(3, 0);
    const FbTable hdr = msg.table(2);
    const uint8_t* body = p + pos + msize;
    if (pos + msize + (size_t)body_len > n) throw std::runtime_error("arrow ipc: truncated message body");
    pos += msize + (size_t)body_len;
    if (htype == MSG_SCHEMA) {
      size_t first;
      const uint32_t nf = hdr.vec(1, &first);
      if (have_schema) {
        if (nf != cols.size()) throw std::runtime_error("arrow ipc: concatenated streams disagree on the schema");
        continue;
      }
      for (uint32_t i = 0; i < nf; i++) {
        const FbTable f = hdr.vec_table(first, i);
        size_t cf;
        if (f.vec(5, &cf) != 0) throw std::runtime_error("arrow ipc: nested column '" + f.str(0) + "'");
        if (f.table(4).ok()) throw std::runtime_error("arrow ipc: dictionary-encoded column '" + f.str(0) + "'");
        HostCol c;
        c.name = f.str(0);
        c.nullable = f.scalar<uint8_t>(1, 0) != 0;
        c.type = read_type(f);
        cols.push_back(c);
      }
      have_schema = true;
    } else if (htype == MSG_RECORD_BATCH) {
      if (!have_schema) throw std::runtime_error("arrow ipc: record batch before schema");
      const int64_t rows = hdr.scalar<int64_t>(0, 0);
      size_t nodes_first, bufs_first;
      const uint32_t n_nodes = hdr.vec(1, &nodes_first);
      const uint32_t n_bufs = hdr.vec(2, &bufs_first);
      const FbTable comp = hdr.table(3);
      const bool compressed = comp.ok();
      if (compressed && comp.scalar<int8_t>(0, 0) != 0) throw std::runtime_error("arrow ipc: body compression codec other than LZ4_FRAME");
      if (n_nodes != cols.size()) throw std::runtime_error("arrow ipc: field node count differs from the schema");
      uint32_t bi = 0;
      std::vector<uint8_t> tmp;
      auto next_buffer = [&](std::vector<uint8_t>& dst, size_t expect_min) {
        if (bi >= n_bufs) throw std::runtime_error("arrow ipc: missing buffer");
        int64_t off, len;
        memcpy(&off, hdr.buf + bufs_first + 16 * (size_t)bi, 8);
        memcpy(&len, hdr.buf + bufs_first + 16 * (size_t)bi + 8, 8);
        bi++;
        if (off < 0 || len < 0 || off + len > body_len) throw std::runtime_error("arrow ipc: buffer outside the body");
        const uint8_t* b = body + off;
        if (len == 0) {
          dst.clear();
        } else if (!compressed) {
          dst.assign(b, b + len);
        } else {
          int64_t ulen;
          memcpy(&ulen, b, 8);
          if (ulen == -1) dst.assign(b + 8, b + len);
          else {
            dst.resize((size_t)ulen);
            lz4_frame_decode(b + 8, (size_t)len - 8, dst.data(), (size_t)ulen);
          }
        }
        if (dst.size() < expect_min) throw std::runtime_error("arrow ipc: buffer shorter than its rows need");
      };
      for (size_t ci = 0; ci < cols.size(); ci++) {
        HostCol& c = cols[ci];
        int64_t node_len, node_nulls;
        memcpy(&node_len, hdr.buf + nodes_first + 16 * ci, 8);
        memcpy(&node_nulls, hdr.buf + nodes_first + 16 * ci + 8, 8);
        if (node_len != rows) throw std::runtime_error("arrow ipc: field node length differs from the batch length");
        if (c.type.id == TypeId::Null) {
          c.n += rows;
          c.null_count += rows;
          continue;
        }
        std::vector<uint8_t> validity, data, extra;
        next_buffer(validity, 0);
        const size_t vbytes = (size_t)((rows + 7) / 8);
        if (c.type.id == TypeId::Bool) next_buffer(data, vbytes);
        else if (c.type.id == TypeId::Utf8) next_buffer(data, (size_t)(rows + 1) * 4 * (rows ? 1 : 0));
        else next_buffer(data, (size_t)rows * (size_t)c.type.width());
        if (c.type.id == TypeId::Utf8) next_buffer(extra, 0);
        // append: bitmaps bit by bit at the running row offset, values byte-wise, offsets rebased
        const int64_t base = c.n;
        auto set_bit = [](std::vector<uint8_t>& bm, int64_t i, bool v) {
          if ((size_t)(i >> 3) >= bm.size()) bm.resize((size_t)(i >> 3) + 1, 0);
          if (v) bm[(size_t)(i >> 3)] |= (uint8_t)(1u << (i & 7));
        };
        const bool has_nulls = node_nulls > 0 && !validity.empty();
        if (has_nulls && c.validity.empty() && base > 0)
          for (int64_t i = 0; i < base; i++) set_bit(c.validity, i, true);
        if (has_nulls || !c.validity.empty())
          for (int64_t i = 0; i < rows; i++) set_bit(c.validity, base + i, has_nulls ? ((validity[(size_t)(i >> 3)] >> (i & 7)) & 1) : true);
        if (has_nulls) c.null_count += node_nulls;
        if (c.type.id == TypeId::Bool) {
          for (int64_t i = 0; i < rows; i++) set_bit(c.data, base + i, (data[(size_t)(i >> 3)] >> (i & 7)) & 1);
          if (c.data.size() < (size_t)((base + rows + 7) / 8)) c.data.resize((size_t)((base + rows + 7) / 8), 0);
        } else if (c.type.id == TypeId::Utf8) {
          if (c.data.empty()) c.data.assign(4, 0);
          const int32_t* off = (const int32_t*)data.data();
          const int32_t first_off = rows ? off[0] : 0;
          const int32_t cur = (int32_t)c.extra.size();
          const size_t old = c.data.size();
          c.data.resize(old + (size_t)rows * 4);
          int32_t* dst = (int32_t*)(c.data.data() + old);
          for (int64_t i = 0; i < rows; i++) dst[i] = cur + (off[i + 1] - first_off);
          if (rows) c.extra.insert(c.extra.end(), extra.begin() + first_off, extra.begin() + off[rows]);
        } else {
          c.data.insert(c.data.end(), data.begin(), data.begin() + (long)((size_t)rows * (size_t)c.type.width()));
        }
        c.n += rows;
      }
      total_rows += rows;
    } else {
      throw std::runtime_error("arrow ipc: message type " + std::to_string(htype) + " (dictionary batches are not supported)");
    }
  }
  for (auto& c : cols) {
    if (c.type.id == TypeId::Utf8 && c.data.empty()) c.data.assign(4, 0);
    if (c.null_count == 0) c.validity.clear();
  }
  return total_rows;
}

}  // namespace ipc
}  // namespace b200
