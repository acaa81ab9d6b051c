// Input pipeline, native part (host code): TFRecord framing + tf.train.Example decoding for the
// 1000-samples-per-record layout of the reference's Criteo files
//   models/wide_deep/src/datasets.py:272-326   (TFRecordDataset with columns feat_ids int32, feat_vals float32,
//                                               label float32; batch = batch_size / line_per_sample records)
//   datasets/criteo_1tb/process_data.py:203-283 (one record = 1000 samples x 39 fields, flattened)
// The reference delegates this to MindSpore's C++ dataset engine; here it is a handful of plain C entry points that the
// Python loader (mindrec_b200/data.py) calls through ctypes on memory-mapped files, decoding straight into pinned host
// buffers that the copy stream then moves to the device.  None of these functions touches the GPU.
//
// TFRecord framing (TensorFlow's RecordWriter): u64 length | u32 masked crc32c(length) | data | u32 masked crc32c(data).
// Example wire format: Example{1: Features{1: map<string, Feature>}}, Feature{1: BytesList | 2: FloatList | 3: Int64List},
// FloatList{1: packed or repeated fixed32}, Int64List{1: packed or repeated varint}.
#include "common.cuh"

namespace mrec {

static uint32_t g_crc_table[8][256];
static bool g_crc_ready = false;

static void crc_init() {
  if (g_crc_ready) return;
  const uint32_t poly = 0x82f63b78u;  // CRC-32C (Castagnoli), reflected
  for (uint32_t i = 0; i < 256; ++i) {
    uint32_t c = i;
    for (int k = 0; k < 8; ++k) c = (c & 1u) ? (c >> 1) ^ poly : (c >> 1);
    g_crc_table[0][i] = c;
  }
  for (uint32_t i = 0; i < 256; ++i)
    for (int t = 1; t < 8; ++t) g_crc_table[t][i] = (g_crc_table[t - 1][i] >> 8) ^ g_crc_table[0][g_crc_table[t - 1][i] & 0xffu];
  g_crc_ready = true;
}

static uint32_t crc32c(const uint8_t* p, size_t n) {
  crc_init();
  uint32_t c = 0xffffffffu;
  while (n >= 8) {  // slicing-by-8
    uint64_t w;
    memcpy(&w, p, 8);
    w ^= c;
    c = g_crc_table[7][w & 0xff] ^ g_crc_table[6][(w >> 8) & 0xff] ^ g_crc_table[5][(w >> 16) & 0xff] ^
        g_crc_table[4][(w >> 24) & 0xff] ^ g_crc_table[3][(w >> 32) & 0xff] ^ g_crc_table[2][(w >> 40) & 0xff] ^
        g_crc_table[1][(w >> 48) & 0xff] ^ g_crc_table[0][(w >> 56) & 0xff];
    p += 8;
    n -= 8;
  }
  while (n--) c = (c >> 8) ^ g_crc_table[0][(c ^ *p++) & 0xffu];
  return c ^ 0xffffffffu;
}

static inline uint32_t mask_crc(uint32_t c) { return ((c >> 15) | (c << 17)) + 0xa282ead8u; }

// protobuf helpers: return false on truncated / malformed input
static inline bool get_varint(const uint8_t*& p, const uint8_t* end, uint64_t* v) {
  uint64_t r = 0;
  for (int shift = 0; shift < 64 && p < end; shift += 7) {
    const uint8_t b = *p++;
    r |= (uint64_t)(b & 0x7f) << shift;
    if (!(b & 0x80)) {
      *v = r;
      return true;
    }
  }
  return false;
}
static inline bool skip_field(const uint8_t*& p, const uint8_t* end, int wire) {
  uint64_t v;
  switch (wire) {
    case 0: return get_varint(p, end, &v);
    case 1: if (end - p < 8) return false; p += 8; return true;
    case 2: if (!get_varint(p, end, &v) || (uint64_t)(end - p) < v) return false; p += v; return true;
    case 5: if (end - p < 4) return false; p += 4; return true;
    default: return false;
  }
}
// enter a length-delimited field: [p, sub_end)
static inline bool get_len(const uint8_t*& p, const uint8_t* end, const uint8_t** sub_end) {
  uint64_t n;
  if (!get_varint(p, end, &n) || (uint64_t)(end - p) < n) return false;
  *sub_end = p + n;
  return true;
}

// kind 0: Int64List -> int32 out; kind 1: FloatList -> float out
static int decode_list(const uint8_t* p, const uint8_t* end, int kind, void* out, int64_t cap, int64_t* count) {
  int64_t n = 0;
  while (p < end) {
    uint64_t tag;
    if (!get_varint(p, end, &tag)) return ERR_SHAPE;
    const int field = (int)(tag >> 3), wire = (int)(tag & 7);
    if (field != 1) {
      if (!skip_field(p, end, wire)) return ERR_SHAPE;
      continue;
    }
    if (kind == 1) {
      if (wire == 2) {  // packed fixed32
        const uint8_t* se;
        if (!get_len(p, end, &se) || ((se - p) & 3)) return ERR_SHAPE;
        const int64_t k = (se - p) / 4;
        if (n + k > cap) return ERR_WORKSPACE;
        memcpy(reinterpret_cast<float*>(out) + n, p, (size_t)k * 4);
        n += k;
        p = se;
      } else if (wire == 5) {
        if (end - p < 4) return ERR_SHAPE;
        if (n + 1 > cap) return ERR_WORKSPACE;
        memcpy(reinterpret_cast<float*>(out) + n, p, 4);
        ++n;
        p += 4;
      } else {
        return ERR_DTYPE;
      }
    } else {
      if (wire == 2) {  // packed varints
        const uint8_t* se;
        if (!get_len(p, end, &se)) return ERR_SHAPE;
        while (p < se) {
          uint64_t v;
          if (!get_varint(p, se, &v)) return ERR_SHAPE;
          if (n + 1 > cap) return ERR_WORKSPACE;
          reinterpret_cast<int32_t*>(out)[n++] = (int32_t)(int64_t)v;
        }
      } else if (wire == 0) {
        uint64_t v;
        if (!get_varint(p, end, &v)) return ERR_SHAPE;
        if (n + 1 > cap) return ERR_WORKSPACE;
        reinterpret_cast<int32_t*>(out)[n++] = (int32_t)(int64_t)v;
      } else {
        return ERR_DTYPE;
      }
    }
  }
  *count = n;
  return OK;
}

}  // namespace mrec

using namespace mrec;

MREC_API uint32_t mrec_crc32c(const void* data, size_t n) { return crc32c(reinterpret_cast<const uint8_t*>(data), n); }
MREC_API uint32_t mrec_crc32c_masked(const void* data, size_t n) {
  return mask_crc(crc32c(reinterpret_cast<const uint8_t*>(data), n));
}

// Scan the framing of a whole file image.  offsets[i] / lengths[i] = payload of record i.  Returns the number of
// records (<= max_records are stored), or -(byte position + 1) of the first framing / CRC error.
MREC_API int64_t mrec_tfrecord_index(const uint8_t* buf, int64_t n, int64_t* offsets, int64_t* lengths, int64_t max_records,
                                     int check_crc) {
  int64_t pos = 0, count = 0;
  while (pos < n) {
    if (n - pos < 12) return -(pos + 1);
    uint64_t len;
    uint32_t lcrc;
    memcpy(&len, buf + pos, 8);
    memcpy(&lcrc, buf + pos + 8, 4);
    if (check_crc && mask_crc(crc32c(buf + pos, 8)) != lcrc) return -(pos + 1);
    if ((uint64_t)(n - pos - 12) < len + 4) return -(pos + 1);
    if (check_crc) {
      uint32_t dcrc;
      memcpy(&dcrc, buf + pos + 12 + len, 4);
      if (mask_crc(crc32c(buf + pos + 12, (size_t)len)) != dcrc) return -(pos + 1);
    }
    if (count < max_records) {
      if (offsets) offsets[count] = pos + 12;
      if (lengths) lengths[count] = (int64_t)len;
    }
    ++count;
    pos += 12 + (int64_t)len + 4;
  }
  return count;
}

// Decode feature `name` of one serialized tf.train.Example.  kind 0: Int64List -> int32 out[cap]; kind 1: FloatList ->
// float out[cap].  *count = values written.  Returns MREC_OK, MREC_ERR_NULL when the feature is absent, MREC_ERR_DTYPE
// when it holds another list type, MREC_ERR_WORKSPACE when out is too small, MREC_ERR_SHAPE on malformed input.
MREC_API int mrec_tfrecord_parse(const uint8_t* rec, int64_t len, const char* name, int kind, void* out, int64_t cap,
                                 int64_t* count) {
  const uint8_t *p = rec, *end = rec + len;
  const size_t name_len = strlen(name);
  *count = 0;
  while (p < end) {  // Example
    uint64_t tag;
    if (!get_varint(p, end, &tag)) return fail(ERR_SHAPE, "mrec_tfrecord_parse: truncated Example");
    if ((tag >> 3) != 1 || (tag & 7) != 2) {
      if (!skip_field(p, end, (int)(tag & 7))) return fail(ERR_SHAPE, "mrec_tfrecord_parse: malformed Example");
      continue;
    }
    const uint8_t* fend;
    if (!get_len(p, end, &fend)) return fail(ERR_SHAPE, "mrec_tfrecord_parse: malformed Features");
    while (p < fend) {  // Features: repeated map entries
      if (!get_varint(p, fend, &tag)) return fail(ERR_SHAPE, "mrec_tfrecord_parse: truncated Features");
      if ((tag >> 3) != 1 || (tag & 7) != 2) {
        if (!skip_field(p, fend, (int)(tag & 7))) return fail(ERR_SHAPE, "mrec_tfrecord_parse: malformed Features");
        continue;
      }
      const uint8_t* eend;
      if (!get_len(p, fend, &eend)) return fail(ERR_SHAPE, "mrec_tfrecord_parse: malformed map entry");
      bool match = false;
      const uint8_t *vbeg = nullptr, *vend = nullptr;
      while (p < eend) {  // map entry: 1 = key, 2 = Feature
        if (!get_varint(p, eend, &tag)) return fail(ERR_SHAPE, "mrec_tfrecord_parse: truncated map entry");
        const int field = (int)(tag >> 3);
        if ((tag & 7) != 2) {
          if (!skip_field(p, eend, (int)(tag & 7))) return fail(ERR_SHAPE, "mrec_tfrecord_parse: malformed map entry");
          continue;
        }
        const uint8_t* se;
        if (!get_len(p, eend, &se)) return fail(ERR_SHAPE, "mrec_tfrecord_parse: malformed map entry");
        if (field == 1) match = ((size_t)(se - p) == name_len) && memcmp(p, name, name_len) == 0;
        else if (field == 2) { vbeg = p; vend = se; }
        p = se;
      }
      if (!match || !vbeg) continue;
      const uint8_t* q = vbeg;  // Feature: oneof kind
      while (q < vend) {
        if (!get_varint(q, vend, &tag)) return fail(ERR_SHAPE, "mrec_tfrecord_parse: truncated Feature");
        const int field = (int)(tag >> 3);
        if ((tag & 7) != 2) {
          if (!skip_field(q, vend, (int)(tag & 7))) return fail(ERR_SHAPE, "mrec_tfrecord_parse: malformed Feature");
          continue;
        }
        const uint8_t* le;
        if (!get_len(q, vend, &le)) return fail(ERR_SHAPE, "mrec_tfrecord_parse: malformed Feature");
        if (field == (kind == 1 ? 2 : 3)) {
          const int rc = decode_list(q, le, kind, out, cap, count);
          if (rc != OK) return fail(rc, "mrec_tfrecord_parse: feature '%s': bad list (code %d)", name, rc);
          return OK;
        }
        if (field >= 1 && field <= 3) return fail(ERR_DTYPE, "mrec_tfrecord_parse: feature '%s' holds another list type", name);
        q = le;
      }
      return fail(ERR_DTYPE, "mrec_tfrecord_parse: feature '%s' is empty", name);
    }
  }
  return fail(ERR_NULL, "mrec_tfrecord_parse: feature '%s' not found", name);
}

// Packed varint encoding of int32 values (writer side of Int64List; negative values take 10 bytes as in protobuf).
// Returns the number of bytes written, or -1 when out is too small.
MREC_API int64_t mrec_varint_pack(const int32_t* v, int64_t n, uint8_t* out, int64_t cap) {
  int64_t o = 0;
  for (int64_t i = 0; i < n; ++i) {
    uint64_t x = (uint64_t)(int64_t)v[i];
    do {
      if (o >= cap) return -1;
      uint8_t b = x & 0x7f;
      x >>= 7;
      out[o++] = b | (x ? 0x80 : 0);
    } while (x);
  }
  return o;
}
