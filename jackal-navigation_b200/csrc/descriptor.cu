// descriptor.cu -- Sobel du/dv + 16-byte descriptor, one fused tiled kernel.
//
// Replaces Descriptor::Descriptor (descriptor.cpp:28-36), filter::sobel3x3
// (filter.cpp:408-416 with 372-405, 227-267, 176-222) and createDescriptor
// (descriptor.cpp:80-112).  Output layout = Descriptor::I_desc: 16 bytes per
// pixel at ((v*W+u)*16); pixels outside u in [3,W-4], v in [3,H-4] are 0
// (the reference leaves them unwritten).
//
// Roofline: HBM.  Algorithmic bytes per image = 1*N read + 16*N written.  The image tile
// (+3 px halo) is staged once in shared memory -- by the TMA unit: one 1-D bulk async copy
// (cp.async.bulk, UBLKCP) per tile row into a 16-byte aligned 96-byte row, completion on an
// mbarrier; inputs whose stride or base is not 16-byte aligned take a plain-load path.  du/dv
// never touch HBM.  The gather works on 32-bit words: a thread builds the descriptors of four
// adjacent pixels from 16 word loads of the du/dv tiles and 30 byte permutes (PRMT), the warp
// transposes them through a swizzled shared-memory buffer, and every store instruction writes
// 32 consecutive descriptors = 512 contiguous bytes.
#include "common.cuh"
#include "blockutil.cuh"

namespace {

constexpr int TW = 64, TH = 16;          // output tile
constexpr int IW = TW + 6, IH = TH + 6;  // image tile with 3-px halo
constexpr int GW_ = TW + 4, GH_ = TH + 4; // gradient tile with 2-px halo
constexpr int IWP = 96;                  // image tile row: x0-16 .. x0+79 (16-byte aligned for bulk copies)
constexpr int IOFF = 13;                 // column of pixel x0-3 inside a tile row
constexpr int GWP = 72;                  // gradient tile row: x0-4 .. x0+67 (word aligned for the gather)

// two 16-bit lanes of a word: h = 0 -> bytes 0 and 2, h = 1 -> bytes 1 and 3
__device__ __forceinline__ uint32_t lanes(uint32_t x, int h) { return (h ? (x >> 8) : x) & 0x00FF00FFu; }
// per 16-bit lane: y = value + 1024 with value in [-1020, 1020]  ->  sat8((value >> 2) + 128)
// (arithmetic shift = floor, filter.cpp:262-264: the reference shifts the 16-bit sums, then packs
// with unsigned saturation)
__device__ __forceinline__ uint32_t sat_lanes(uint32_t y) {
  const uint32_t z = (y >> 2) & 0x01FF01FFu;                       // (value >> 2) + 256
  return __vsub2(__vminu2(__vmaxu2(z, 0x00800080u), 0x017F017Fu), 0x00800080u);
}

__global__ void __launch_bounds__(256) descriptor_kernel(Geo g, const uint8_t* __restrict__ I1,
                                                         const uint8_t* __restrict__ I2,
                                                         uint8_t* __restrict__ D1, uint8_t* __restrict__ D2,
                                                         int tma_ok) {
  __shared__ __align__(128) uint8_t sIraw[IH * IWP];
  __shared__ __align__(16) uint8_t sU[GH_ * GWP];
  __shared__ __align__(16) uint8_t sV[GH_ * GWP];
  __shared__ uint4 sOut[8][2 * TW];       // per warp: two rows of descriptors, swizzled
  __shared__ __align__(8) uint64_t s_bar;
  uint8_t* sI = sIraw + IOFF;   // sI[r * IWP + c] = pixel (x0 - 3 + c, y0 - 3 + r)

  const int W = g.W, H = g.H;
  const int frame = blockIdx.z >> 1, side = blockIdx.z & 1;
  const uint8_t* I = (side ? I2 : I1) + (size_t)frame * g.bpl * H;
  uint8_t* D = (side ? D2 : D1) + (size_t)frame * W * H * 16;
  const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
  const int tid = threadIdx.x;

  // 1. image tile.  Pixels outside the image never reach a descriptor that is kept, so only the
  //    bytes no copy writes are cleared.
  if (tma_ok) {
    const int xs = max(x0 - 16, 0), xe = min(x0 + 80, g.bpl & ~15);   // copied byte range of a row
    const int r0 = max(0, 3 - y0), r1 = min(IH, H + 3 - y0);          // tile rows inside the image
    if (tid == 0) {
      mbar_init(&s_bar, 1);
      fence_mbar_init();
    }
    // interior tiles are overwritten completely by the copies; only tiles that touch the image
    // border have bytes no copy writes
    if (r0 > 0 || r1 < IH || xs > x0 - 16 || xe < x0 + 80)
      for (int i = tid; i < IH * IWP; i += 256) {
        const int r = i / IWP, x = x0 - 16 + (i - r * IWP);
        if (r < r0 || r >= r1 || x < xs || x >= xe) sIraw[i] = 0;
      }
    __syncthreads();
    if (tid == 0) {
      mbar_arrive_expect_tx(&s_bar, (uint32_t)((r1 - r0) * (xe - xs)));
      for (int r = r0; r < r1; r++)
        bulk_g2s(sIraw + r * IWP + (xs - (x0 - 16)), I + (size_t)(y0 - 3 + r) * g.bpl + xs, (uint32_t)(xe - xs),
                 &s_bar);
      mbar_wait(&s_bar, 0);   // one waiting thread; the others sleep in the barrier instead of polling
    }
    __syncthreads();
  } else {
    for (int i = tid; i < IH * IW; i += 256) {
      int r = i / IW, c = i - r * IW;
      int y = y0 - 3 + r, x = x0 - 3 + c;
      uint8_t v = 0;
      if (x >= 0 && x < W && y >= 0 && y < H) v = __ldg(I + (size_t)y * g.bpl + x);
      sI[r * IWP + c] = v;
    }
    __syncthreads();
  }

  // 2. Sobel responses, four gradient pixels (one word of a gradient tile row) per thread, two
  //    16-bit lanes per register: E lanes = bytes 0 and 2 of a word, O lanes = bytes 1 and 3.
  //    Tile word w of row r = gradient columns x0-4+4w .. +3 of image row y0-2+r; it needs image
  //    columns -1 .. +4 around them = bytes 4w+11 .. 4w+16 of the image tile rows r, r+1, r+2.
  {
    const uint32_t* sI32 = reinterpret_cast<const uint32_t*>(sIraw);
    uint32_t* sU32 = reinterpret_cast<uint32_t*>(sU);
    uint32_t* sV32 = reinterpret_cast<uint32_t*>(sV);
    constexpr int GW4 = GWP / 4, IW4 = IWP / 4;
    for (int i = tid; i < GH_ * GW4; i += 256) {
      const int r = i / GW4, w = i - r * GW4;
      const uint32_t* p = sI32 + r * IW4 + w + 2;
      const uint32_t A0 = p[0], A1 = p[1], A2 = p[2];
      const uint32_t B0 = p[IW4], B1 = p[IW4 + 1], B2 = p[IW4 + 2];
      const uint32_t C0 = p[2 * IW4], C1 = p[2 * IW4 + 1], C2 = p[2 * IW4 + 2];
      // columns -1..+2 (left neighbours), 0..+3 (centres), +1..+4 (right neighbours) of the 4 pixels
      const uint32_t al = __funnelshift_r(A0, A1, 24), ar = __funnelshift_r(A1, A2, 8);
      const uint32_t bl = __funnelshift_r(B0, B1, 24), br = __funnelshift_r(B1, B2, 8);
      const uint32_t cl = __funnelshift_r(C0, C1, 24), cr = __funnelshift_r(C1, C2, 8);
      uint32_t du[2], dv[2];
#pragma unroll
      for (int h = 0; h < 2; h++) {   // h = 0: E lanes (pixels 0, 2), h = 1: O lanes (pixels 1, 3)
        const uint32_t ale = lanes(al, h), are = lanes(ar, h), ame = lanes(A1, h);
        const uint32_t ble = lanes(bl, h), bre = lanes(br, h);
        const uint32_t cle = lanes(cl, h), cre = lanes(cr, h), cme = lanes(C1, h);
        // S = above + 2 centre + below (0..1020), T + 255 = above - below + 255 (0..510)
        const uint32_t Sl = ale + cle + 2u * ble, Sr = are + cre + 2u * bre;
        const uint32_t Tl = ale + (cle ^ 0x00FF00FFu), Tm = ame + (cme ^ 0x00FF00FFu), Tr = are + (cre ^ 0x00FF00FFu);
        du[h] = sat_lanes(Sl + 0x04000400u - Sr);             // S(u-1) - S(u+1) + 1024
        dv[h] = sat_lanes(Tl + 2u * Tm + Tr + 0x00040004u);   // T(u-1) + 2 T(u) + T(u+1) + 1024
      }
      sU32[i] = du[0] | (du[1] << 8);
      sV32[i] = dv[0] | (dv[1] << 8);
    }
  }
  __syncthreads();

  // 3. gather.  Thread (q, ly) builds pixels x0+4q .. x0+4q+3 of row y0+ly.  In a tile row, word
  //    q holds gradient columns x0+4q-4 .. x0+4q-1: L/M/R = words q, q+1, q+2 form a 12-byte
  //    window whose byte 4+j+o is column (pixel j) + o.
  const int q = tid & 15, ly = tid >> 4, warp = tid >> 5, lane = tid & 31;
  {
    const uint32_t* u32 = reinterpret_cast<const uint32_t*>(sU) + ly * (GWP / 4) + q;
    const uint32_t* v32 = reinterpret_cast<const uint32_t*>(sV) + ly * (GWP / 4) + q;
    constexpr int RW = GWP / 4;   // words per tile row; output row ly = tile rows ly .. ly+4
    const uint32_t um2 = u32[1];
    const uint32_t a0 = u32[RW], a1 = u32[RW + 1], a2 = u32[RW + 2];              // du row -1
    const uint32_t b0 = u32[2 * RW], b1 = u32[2 * RW + 1], b2 = u32[2 * RW + 2];  // du row  0
    const uint32_t c0 = u32[3 * RW], c1 = u32[3 * RW + 1], c2 = u32[3 * RW + 2];  // du row +1
    const uint32_t up2 = u32[4 * RW + 1];
    const uint32_t vm1 = v32[RW + 1], vp1 = v32[3 * RW + 1];
    const uint32_t d0 = v32[2 * RW], d1 = v32[2 * RW + 1], d2 = v32[2 * RW + 2];  // dv row 0
    uint4 o[4];
    // bytes 0..3 of each word, in descriptor order (descriptor.cpp:84-100)
    o[0].x = __byte_perm(__byte_perm(a0, a1, 0x6420), um2, 0x3214);
    o[1].x = __byte_perm(__byte_perm(a0, a1, 0x7530), um2, 0x3215);
    o[2].x = __byte_perm(__byte_perm(a1, a2, 0x4200), um2, 0x3216);
    o[3].x = __byte_perm(__byte_perm(a1, a2, 0x5310), um2, 0x3217);
    o[0].y = __byte_perm(b0, b1, 0x5443);
    o[1].y = __byte_perm(b1, b1, 0x2110);
    o[2].y = __byte_perm(b1, b1, 0x3221);
    o[3].y = __byte_perm(b1, b2, 0x4332);
    o[0].z = __byte_perm(__byte_perm(c0, c1, 0x0642), up2, 0x4210);
    o[1].z = __byte_perm(__byte_perm(c0, c1, 0x0753), up2, 0x5210);
    o[2].z = __byte_perm(__byte_perm(c1, c2, 0x0420), up2, 0x6210);
    o[3].z = __byte_perm(__byte_perm(c1, c2, 0x0531), up2, 0x7210);
    o[0].w = __byte_perm(__byte_perm(vm1, vp1, 0x4000), __byte_perm(d0, d1, 0x0530), 0x3650);
    o[1].w = __byte_perm(__byte_perm(vm1, vp1, 0x5001), d1, 0x3640);
    o[2].w = __byte_perm(__byte_perm(vm1, vp1, 0x6002), d1, 0x3750);
    o[3].w = __byte_perm(__byte_perm(vm1, vp1, 0x7003), __byte_perm(d1, d2, 0x0420), 0x3650);
    // pixel p (0..127: two rows of 64) of this warp -> slot; the XOR keeps both the 64-byte strided
    // writes here and the linear reads below free of bank conflicts
    const int pq = (lane >> 4) * 16 + q;   // quad index inside the warp's two rows
#pragma unroll
    for (int j = 0; j < 4; j++) sOut[warp][4 * pq + (j ^ ((pq >> 1) & 3))] = o[j];
  }
  __syncwarp();
  const bool sub = g.p.subsampling != 0;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int p = 32 * i + lane, pq = p >> 2;
    const int x = x0 + (p & 63), y = y0 + 2 * warp + (p >> 6);
    if (x >= W || y >= H) continue;
    uint4 out = sOut[warp][4 * pq + ((p & 3) ^ ((pq >> 1) & 3))];
    // half resolution (descriptor.cpp:48-78): only rows 4, 6, 8, ... carry descriptors
    const bool row_ok = sub ? (y >= 4 && !(y & 1)) : (y >= 3);
    if (!(x >= 3 && x <= W - 4 && row_ok && y <= H - 4)) out = make_uint4(0, 0, 0, 0);
    *reinterpret_cast<uint4*>(D + ((size_t)y * W + x) * 16) = out;
  }
}

}  // namespace

void launch_descriptor(const Geo& g, int B, const uint8_t* I1, const uint8_t* I2, Workspace& ws, cudaStream_t s) {
  dim3 grid((g.W + TW - 1) / TW, (g.H + TH - 1) / TH, 2 * B);
  // bulk async copies need 16-byte aligned global addresses and sizes
  const int tma_ok = (g.bpl % 16 == 0) && ((((uintptr_t)I1) | ((uintptr_t)I2)) % 16 == 0) &&
                     (((size_t)g.bpl * g.H) % 16 == 0);
  descriptor_kernel<<<grid, 256, 0, s>>>(g, I1, I2, ws.desc[0], ws.desc[1], tma_ok);
  g_jn_launches += 1;
}
