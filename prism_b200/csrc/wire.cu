// wire.cu -- host-side codec of the remote-actor transport.  No device code.
//
// The reference moves collector steps and training batches through Redis as msgpack arrays of Python numbers
// (prism/async_components/compression_methods.py:59-70, redis/redis_interface.py:113-137,
// async_experience_buffer.py:84-96) and walks them element by element in Python
// (prism/experience/timestep.py:103-187).  Here a payload is decoded straight into a float64 array, a block of
// Timestep records is indexed in one pass (offsets only: the observations are then sliced as array views), and a
// batch is encoded from the tensors' memory -- byte-identical to what msgpack emits for the same Python lists.
#include <string.h>
#include "common.cuh"

namespace {

constexpr double WIRE_NULL = -1313.0;             // NULL_VALUE, prism/experience/timestep.py:9
constexpr int REC_W = 12;                         // columns of one indexed record, see pb_wire_index_timesteps
constexpr int TAIL_FIELDS = 12;                   // reward .. next id, timestep.py:58-96

inline uint64_t be(const unsigned char *p, int n)
{
    uint64_t v = 0;
    for (int i = 0; i < n; ++i) v = (v << 8) | p[i];
    return v;
}

inline void put_be(unsigned char *p, uint64_t v, int n)
{
    for (int i = n - 1; i >= 0; --i) { p[i] = (unsigned char)(v & 0xff); v >>= 8; }
}

}  // namespace

extern "C" {

int pb_wire_unpack_numbers(const unsigned char *buf, long long len, double *out, long long cap, long long *n_out)
{
    if (!buf || len < 1 || !n_out) return PB_E_ARG;
    long long pos = 0, n = 0;
    const unsigned char h = buf[pos++];
    if ((h & 0xf0) == 0x90) n = h & 0x0f;
    else if (h == 0xdc) { if (len < pos + 2) return PB_E_ARG; n = (long long)be(buf + pos, 2); pos += 2; }
    else if (h == 0xdd) { if (len < pos + 4) return PB_E_ARG; n = (long long)be(buf + pos, 4); pos += 4; }
    else return PB_E_UNSUPPORTED;                                   // not an array
    *n_out = n;
    if (out && cap < n) return PB_E_ARG;
    for (long long i = 0; i < n; ++i) {
        if (pos >= len) return PB_E_ARG;
        const unsigned char t = buf[pos++];
        double v;
        int w = 0;
        if (t <= 0x7f) v = (double)t;                               // positive fixint
        else if (t >= 0xe0) v = (double)(int8_t)t;                  // negative fixint
        else {
            switch (t) {
                case 0xc2: v = 0.0; break;
                case 0xc3: v = 1.0; break;
                case 0xca: w = 4; break;
                case 0xcb: w = 8; break;
                case 0xcc: case 0xd0: w = 1; break;
                case 0xcd: case 0xd1: w = 2; break;
                case 0xce: case 0xd2: w = 4; break;
                case 0xcf: case 0xd3: w = 8; break;
                default: return PB_E_UNSUPPORTED;                   // nil, str, bin, nested containers
            }
            if (w) {
                if (pos + w > len) return PB_E_ARG;
                const uint64_t raw = be(buf + pos, w);
                pos += w;
                if (t == 0xca) { uint32_t r32 = (uint32_t)raw; float f; memcpy(&f, &r32, 4); v = (double)f; }
                else if (t == 0xcb) memcpy(&v, &raw, 8);
                else if (t >= 0xcc && t <= 0xcf) v = (double)raw;
                else if (t == 0xd0) v = (double)(int8_t)raw;
                else if (t == 0xd1) v = (double)(int16_t)raw;
                else if (t == 0xd2) v = (double)(int32_t)raw;
                else v = (double)(int64_t)raw;
            }
        }
        if (out) out[i] = v;
    }
    return pos == len ? PB_OK : PB_E_ARG;                           // trailing bytes: not ONE array
}

int pb_wire_array_header(long long n, unsigned char *out, long long *written)
{
    if (n < 0 || n > 0xffffffffLL || !out || !written) return PB_E_ARG;
    if (n < 16) { out[0] = (unsigned char)(0x90 | n); *written = 1; }
    else if (n < 65536) { out[0] = 0xdc; put_be(out + 1, (uint64_t)n, 2); *written = 3; }
    else { out[0] = 0xdd; put_be(out + 1, (uint64_t)n, 4); *written = 5; }
    return PB_OK;
}

int pb_wire_pack_numbers(const void *src, int dtype, long long n, unsigned char *out, long long cap, long long *written)
{
    if ((!src && n > 0) || n < 0 || !out || !written || dtype < 0 || dtype > 3) return PB_E_ARG;
    if (cap < 9 * n) return PB_E_ARG;                               // worst case: tag + 8 bytes per element
    long long pos = 0;
    for (long long i = 0; i < n; ++i) {
        if (dtype == 0 || dtype == 1) {                             // Python float: always float64 on the wire
            const double v = dtype == 0 ? (double)((const float *)src)[i] : ((const double *)src)[i];
            uint64_t raw;
            memcpy(&raw, &v, 8);
            out[pos++] = 0xcb;
            put_be(out + pos, raw, 8);
            pos += 8;
        } else if (dtype == 3) {
            out[pos++] = ((const unsigned char *)src)[i] ? 0xc3 : 0xc2;
        } else {                                                    // Python int: shortest form
            const int64_t v = ((const int64_t *)src)[i];
            if (v >= 0) {
                if (v < 128) out[pos++] = (unsigned char)v;
                else if (v < 256) { out[pos++] = 0xcc; out[pos++] = (unsigned char)v; }
                else if (v < 65536) { out[pos++] = 0xcd; put_be(out + pos, (uint64_t)v, 2); pos += 2; }
                else if (v < 4294967296LL) { out[pos++] = 0xce; put_be(out + pos, (uint64_t)v, 4); pos += 4; }
                else { out[pos++] = 0xcf; put_be(out + pos, (uint64_t)v, 8); pos += 8; }
            } else {
                if (v >= -32) out[pos++] = (unsigned char)(int8_t)v;
                else if (v >= -128) { out[pos++] = 0xd0; out[pos++] = (unsigned char)(int8_t)v; }
                else if (v >= -32768) { out[pos++] = 0xd1; put_be(out + pos, (uint64_t)v, 2); pos += 2; }
                else if (v >= -2147483648LL) { out[pos++] = 0xd2; put_be(out + pos, (uint64_t)v, 4); pos += 4; }
                else { out[pos++] = 0xd3; put_be(out + pos, (uint64_t)v, 8); pos += 8; }
            }
        }
    }
    *written = pos;
    return PB_OK;
}

// One pass over a flat block of serialized Timesteps (layout: timestep.py:30-101).  Per record, REC_W int64 columns:
//   0 id | 1 obs offset (-1: none) | 2 obs values | 3 shape offset | 4 shape values |
//   5 id of the truncated successor (-1313: none) | 6..9 the same four columns for its observation |
//   10 offset of the 12 scalar fields (reward, done, truncated, action, n_step_return, n_step_gamma, n_step_done,
//      needs_n_step, episodic_reward, n_step_next id, prev id, next id) | 11 offset one past the record.
// rec == NULL counts the records only.
int pb_wire_index_timesteps(const double *flat, long long n, long long max_records, long long *rec, long long *n_rec)
{
    if ((!flat && n > 0) || n < 0 || !n_rec) return PB_E_ARG;
    long long idx = 0, r = 0;
    // {offset, values, shape offset, shape values} of one observation starting at its value count
    auto parse_obs = [&](long long *o) -> bool {
        if (idx >= n) return false;
        const double n_obs = flat[idx++];
        if (!(n_obs >= 0) || n_obs > (double)(n - idx)) return false;
        o[0] = idx; o[1] = (long long)n_obs;
        idx += o[1];
        if (idx >= n) return false;
        const double n_shape = flat[idx++];
        if (!(n_shape >= 0) || n_shape > (double)(n - idx)) return false;
        o[2] = idx; o[3] = (long long)n_shape;
        idx += o[3];
        return true;
    };
    while (idx < n) {
        long long col[REC_W];
        col[0] = (long long)flat[idx++];
        if (idx >= n) return PB_E_ARG;
        if (flat[idx] == WIRE_NULL) { col[1] = -1; col[2] = col[3] = col[4] = 0; ++idx; }
        else if (!parse_obs(col + 1)) return PB_E_ARG;
        if (idx >= n) return PB_E_ARG;
        if (flat[idx] == WIRE_NULL) { col[5] = (long long)WIRE_NULL; col[6] = -1; col[7] = col[8] = col[9] = 0; ++idx; }
        else { col[5] = (long long)flat[idx++]; if (!parse_obs(col + 6)) return PB_E_ARG; }
        if (idx + TAIL_FIELDS > n) return PB_E_ARG;
        col[10] = idx;
        idx += TAIL_FIELDS;
        col[11] = idx;
        if (rec) {
            if (r >= max_records) return PB_E_ARG;
            memcpy(rec + r * REC_W, col, sizeof(col));
        }
        ++r;
    }
    *n_rec = r;
    return PB_OK;
}

}  // extern "C"
