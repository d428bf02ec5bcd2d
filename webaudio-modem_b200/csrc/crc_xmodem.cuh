// crc_xmodem.cuh — warp-level CRC-16/CCITT-FALSE and XModem framed-block check.
//
// CRC16.calculate (src/utils/crc16.ts:21-38): poly 0x1021, init 0xFFFF, no xorout, MSB first.
// Packet layout (src/transports/xmodem/packet.ts:42-54): SOH SEQ ~SEQ LEN payload CRChi CRClo, CRC
// over the payload only.  Receive-side rules (src/transports/xmodem/xmodem.ts:232-321).
//
// One warp owns one byte block.  The CRC is linear over GF(2): each lane reduces a contiguous
// slice of the block from a zero register (byte-wise, 256-entry table in shared memory), the slice
// remainders are advanced to the end of the block by multiplying with x^(8*bytes_after) mod P
// (table of x^(8k) for k < 512 in shared memory, square-and-multiply beyond), and the 32 partial
// remainders are XOR-reduced with warp shuffles.  The 0xFFFF init value is one more term:
// 0xFFFF * x^(8*len) mod P.
#pragma once

#include "wam_common.cuh"

namespace wam {

struct PktResultDev {  // same layout as wam_pkt_result
  int32_t status, sequence, length, payloadOffset, crcReceived, crcComputed, bytesConsumed;
};

constexpr int kCrcPowTable = 512;

struct CrcTables {  // shared memory, filled by every CTA
  uint16_t byte_tab[256];          // remainder of (b << 8) * x^8 ... i.e. classic MSB-first table
  uint16_t pow8[kCrcPowTable];     // x^(8k) mod P
};

__host__ __device__ inline uint32_t crc16_update_byte(uint32_t crc, uint32_t byte) {
  crc ^= byte << 8;
  for (int i = 0; i < 8; ++i) crc = (crc & 0x8000u) ? ((crc << 1) ^ 0x1021u) & 0xffffu : (crc << 1) & 0xffffu;
  return crc;
}

// (a * b) mod P over GF(2), P = x^16 + x^12 + x^5 + 1, operands are 16-bit remainders
__device__ __forceinline__ uint32_t gf_mulmod(uint32_t a, uint32_t b) {
  uint32_t r = 0;
#pragma unroll
  for (int i = 15; i >= 0; --i) {
    r = (r & 0x8000u) ? ((r << 1) ^ 0x1021u) & 0xffffu : (r << 1) & 0xffffu;
    if ((b >> i) & 1u) r ^= a;
  }
  return r;
}

__device__ __forceinline__ void crc_tables_init(CrcTables& t) {
  for (int i = threadIdx.x; i < 256; i += blockDim.x) t.byte_tab[i] = (uint16_t)crc16_update_byte(0u, (uint32_t)i);
  // x^(8k): x^0 = 1, then multiply by x^8 = 0x0100; each thread builds a strided subsequence
  for (int k = threadIdx.x; k < kCrcPowTable; k += blockDim.x) {
    uint32_t result = 1u, base = 0x0100u, e = (uint32_t)k;
    while (e) {
      if (e & 1u) result = gf_mulmod(result, base);
      base = gf_mulmod(base, base);
      e >>= 1;
    }
    t.pow8[k] = (uint16_t)result;
  }
  __syncthreads();
}

__device__ __forceinline__ uint32_t gf_x_pow8(const CrcTables& t, uint32_t k) {
  if (k < (uint32_t)kCrcPowTable) return t.pow8[k];
  uint32_t result = t.pow8[k & (kCrcPowTable - 1)];
  uint32_t base = gf_mulmod(t.pow8[kCrcPowTable - 1], 0x0100u);  // x^(8*512)
  k >>= 9;
  while (k) {
    if (k & 1u) result = gf_mulmod(result, base);
    base = gf_mulmod(base, base);
    k >>= 1;
  }
  return result;
}

// CRC of bytes[0..len) computed by the whole warp; every lane returns the result.
__device__ __forceinline__ uint32_t warp_crc16(const CrcTables& t, const uint8_t* __restrict__ bytes, int len, int lane) {
  const int per = (len + 31) >> 5;
  const int b0 = min(lane * per, len);
  const int b1 = min(b0 + per, len);
  uint32_t part = 0;
  for (int i = b0; i < b1; ++i) part = ((part << 8) & 0xffffu) ^ t.byte_tab[((part >> 8) ^ bytes[i]) & 0xffu];
  if (b1 > b0) part = gf_mulmod(part, gf_x_pow8(t, (uint32_t)(len - b1)));
  if (lane == 0) part ^= gf_mulmod(0xffffu, gf_x_pow8(t, (uint32_t)len));  // init value term
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part ^= __shfl_xor_sync(0xffffffffu, part, o);
  return part & 0xffffu;
}

// one warp per block
__global__ void __launch_bounds__(128) crc16_batch_kernel(const uint8_t* __restrict__ bytes, long stride,
                                                          const int32_t* __restrict__ len, long n_blocks,
                                                          uint16_t* __restrict__ crc_out) {
  __shared__ CrcTables tabs;
  crc_tables_init(tabs);
  const long w = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= n_blocks) return;
  const uint32_t c = warp_crc16(tabs, bytes + w * stride, len[w], lane);
  if (lane == 0) crc_out[w] = (uint16_t)c;
}

// One warp per stream of demodulated bytes.  Each CTA (4 warps) walks a strided set of streams so the
// table set-up is amortised.
__global__ void __launch_bounds__(128) xmodem_check_kernel(const uint8_t* __restrict__ bytes, long stride,
                                                           const int32_t* __restrict__ len,
                                                           const int32_t* __restrict__ expected_seq, long n_streams,
                                                           PktResultDev* __restrict__ res) {
  __shared__ CrcTables tabs;
  crc_tables_init(tabs);
  const int lane = threadIdx.x & 31;
  const long warps_total = (long)gridDim.x * (blockDim.x >> 5);
  for (long w = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < n_streams; w += warps_total) {
    const uint8_t* row = bytes + w * stride;
    const int n = len[w];
    const int expected = expected_seq ? expected_seq[w] : 1;

    PktResultDev r;
    r.status = 2; r.sequence = -1; r.length = -1; r.payloadOffset = -1; r.crcReceived = -1; r.crcComputed = -1;
    r.bytesConsumed = n;

    // receiveAllPackets (xmodem.ts:236-252): the first byte that is SOH or EOT decides
    int first = -1, first_val = 0;
    for (int base = 0; base < n; base += 32) {
      const int i = base + lane;
      const int v = (i < n) ? row[i] : 0;
      const unsigned m = __ballot_sync(0xffffffffu, i < n && (v == 0x01 || v == 0x04));
      if (m) {
        const int src = __ffs(m) - 1;
        first = base + src;
        first_val = __shfl_sync(0xffffffffu, v, src);
        break;
      }
    }
    if (first < 0) {
      r.status = 2;  // NO_SOH
    } else if (first_val == 0x04) {
      r.status = 3;  // EOT
      r.bytesConsumed = first + 1;
    } else {
      int p = first + 1;
      if (p + 3 > n) {
        r.status = 4; r.bytesConsumed = p;
      } else {
        const int seq = row[p], nseq = row[p + 1], plen = row[p + 2];
        p += 3;
        r.sequence = seq; r.length = plen;
        const int prev = expected == 1 ? 255 : expected - 1;  // xmodem.ts:525-530
        if (seq + nseq != 255) {
          r.status = 5; r.bytesConsumed = p;
        } else if (seq == expected) {
          if (p + plen + 2 > n) {
            r.status = 4; r.bytesConsumed = p;
          } else {
            r.payloadOffset = p;
            r.crcReceived = (row[p + plen] << 8) | row[p + plen + 1];
            r.crcComputed = (int)warp_crc16(tabs, row + p, plen, lane);
            r.bytesConsumed = p + plen + 2;
            r.status = (r.crcComputed != r.crcReceived) ? 6 : 0;
          }
        } else if (seq == prev) {
          if (p + plen + 2 > n) {
            r.status = 4; r.bytesConsumed = p;
          } else {
            r.payloadOffset = p;
            r.bytesConsumed = p + plen + 2;
            r.status = 1;
          }
        } else {
          r.status = 7; r.bytesConsumed = p;
        }
      }
    }
    if (lane == 0) res[w] = r;
  }
}

struct XmodemRxStateDev {  // same layout as wam_xmodem_rx_state
  int32_t expectedSequence, retries, done, dataLen, packetsReceived, packetsDropped;
};

// Receive side of XModemTransport.receiveAllPackets / receiveAndProcessPacket (xmodem.ts:232-321) for one
// burst of demodulated bytes per stream, receiver state carried between bursts.  One warp per stream: the
// walk over packets is sequential, but inside it the scan for SOH / EOT is a ballot over 32 bytes at a time,
// the CRC is the warp-level one above and the payload copy is lane-parallel.
__global__ void __launch_bounds__(128) xmodem_receive_kernel(const uint8_t* __restrict__ bytes, long stride,
                                                             const int32_t* __restrict__ len, long n_streams,
                                                             int max_retries, XmodemRxStateDev* __restrict__ state,
                                                             uint8_t* __restrict__ replies, int reply_cap,
                                                             int32_t* __restrict__ n_replies,
                                                             int32_t* __restrict__ consumed,
                                                             uint8_t* __restrict__ data, long data_stride) {
  __shared__ CrcTables tabs;
  crc_tables_init(tabs);
  const int lane = threadIdx.x & 31;
  const long warps_total = (long)gridDim.x * (blockDim.x >> 5);
  for (long w = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < n_streams; w += warps_total) {
    const uint8_t* row = bytes + w * stride;
    const int n = len[w];
    XmodemRxStateDev st = state[w];
    uint8_t* rep = replies + w * (long)reply_cap;
    uint8_t* out = data ? data + w * data_stride : nullptr;
    int p = 0, nrep = 0;
    while (!st.done && p < n) {
      // next byte that is SOH or EOT; everything before it is ignored (xmodem.ts:248-250)
      int first = -1, first_val = 0;
      for (int base = p; base < n; base += 32) {
        const int i = base + lane;
        const int v = (i < n) ? row[i] : 0;
        const unsigned m = __ballot_sync(0xffffffffu, i < n && (v == 0x01 || v == 0x04));
        if (m) {
          const int src = __ffs(m) - 1;
          first = base + src;
          first_val = __shfl_sync(0xffffffffu, v, src);
          break;
        }
      }
      if (first < 0) { p = n; break; }
      p = first;
      if (first_val == 0x04) {  // EOT: final ACK (xmodem.ts:241-244)
        p++;
        if (lane == 0 && nrep < reply_cap) rep[nrep] = 0x06;
        nrep++;
        st.done = 1;
        break;
      }
      if (p + 4 > n) break;  // header not complete yet: keep the SOH
      const int seq = row[p + 1], nseq = row[p + 2], plen = row[p + 3];
      bool error = false;
      if (seq + nseq != 255) {
        st.packetsDropped++;
        error = true;
      } else {
        const int prev = st.expectedSequence == 1 ? 255 : st.expectedSequence - 1;  // xmodem.ts:525-530
        if (seq == st.expectedSequence) {
          if (p + 4 + plen + 2 > n) break;  // payload not complete yet
          st.packetsReceived++;
          const int crc = (row[p + 4 + plen] << 8) | row[p + 4 + plen + 1];
          if ((int)warp_crc16(tabs, row + p + 4, plen, lane) != crc) {
            st.packetsDropped++;
            error = true;
          } else {
            if (out)
              for (int i = lane; i < plen; i += 32)
                if ((long)st.dataLen + i < data_stride) out[st.dataLen + i] = row[p + 4 + i];
            st.dataLen += plen;
            st.expectedSequence = (st.expectedSequence % 255) + 1;
            st.retries = 0;
            if (lane == 0 && nrep < reply_cap) rep[nrep] = 0x06;
            nrep++;
            p += 4 + plen + 2;
          }
        } else if (seq == prev) {  // duplicate: consumed, ACKed, ignored (xmodem.ts:309-314)
          if (p + 4 + plen + 2 > n) break;
          st.packetsDropped++;
          if (lane == 0 && nrep < reply_cap) rep[nrep] = 0x06;
          nrep++;
          p += 4 + plen + 2;
        } else {
          st.packetsDropped++;
          error = true;
        }
      }
      if (error) {  // catch block, xmodem.ts:251-260
        p = n;      // receive.buffer = []
        if (++st.retries > max_retries) { st.done = 2; break; }
        if (lane == 0 && nrep < reply_cap) rep[nrep] = 0x15;
        nrep++;
      }
    }
    if (lane == 0) {
      state[w] = st;
      n_replies[w] = nrep;
      consumed[w] = p;
    }
  }
}

}  // namespace wam
