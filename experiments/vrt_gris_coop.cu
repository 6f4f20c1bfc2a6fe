// EXPERIMENT (round 2, measured and lost): warp-cooperative form of k_gris. Not compiled into libvoxelrt.so.
// Parity-green (frames agree with k_gris to 5e-7, the ReSTIR / GRIS GPU tests pass with it) but 20-24 % SLOWER:
// example6 3.79 -> 4.55 ms, example3 4.11 -> 5.08 ms at 1080p (profiles/r04b_ab_gris_coop.log). A shift() is only
// ~500 warp instructions once its early-outs are counted (k_gris issues ~34 K warp instructions per 8x4 tile for
// 64 tap passes), so the per-pair cost of dealing pairs to lanes (prefix search, __fns, the centre's record read
// back from shared memory, material rows / bases / view vector recomputed per pair instead of once per pixel, the
// G-buffer record loaded twice) is larger than what the idle lanes of rejected taps cost.
// To try it again: paste the kernel below into vrt_restir.cu after k_gris (it uses that file's helpers) and the
// launcher fragment at the end into vrt_launch_gris.

// ---------------------------------------------------------------------------- k_gris_coop
// Warp-cooperative form of k_gris: same arithmetic per (centre pixel, tap) pair, but the pairs that pass the similarity
// test (pathtracer.py:911; about half of them outdoors) are COMPACTED across the warp before the two shifts run, so
// that a rejected tap no longer idles its lane for a whole shift(). Per warp = 8x4 pixel tile, taps in two chunks of 16:
//   accept  every lane tests its own 16 taps (cheap, all lanes busy); __ballot_sync gives one 32-bit mask of accepting
//           centres per tap, kept in shared memory with its running prefix count; the tap offset goes to offM[tap][centre]
//   pass A  the accepted pairs, enumerated tap-major (neighbouring centres of ONE tap -> neighbouring loads, as in
//           k_gris), are dealt 32 at a time to the lanes: lane L of a round takes pair e = 32 * round + L, finds its
//           tap by a 4-step search of the prefix counts and its centre with __fns, reads the centre's sample from the
//           warp's shared-memory records and shifts it to the tap: val[0][tap][centre] = p_hat(centre -> tap)
//   pass B  same enumeration: the tap's reservoir shifted to the centre; the merge weight, 1 - canonical weight and the
//           tap's M go back to shared memory
//   merge   every lane walks its own accepted taps in tap order (reservoir.py:76-86: running sums + one RNG draw per
//           tap) and remembers only WHICH tap it selected
// After the second chunk one more pass-B round re-shifts the selected tap of every centre (one pair per lane) to get
// the integrand and sample that the in-order merge of k_gris would have kept. The chunk loop is not unrolled, so the
// kernel still holds exactly one copy of each shift direction.
#ifndef VRT_GRIS_COOP_DEFAULT
#define VRT_GRIS_COOP_DEFAULT 0  // VRT_GRIS_COOP=1 in the environment selects k_gris_coop
#endif
#define GC_CHUNK 16
#define GC_REC_WORDS 27
#define GC_WARP_WORDS (GC_REC_WORDS * 32 + GC_CHUNK * 32 + 2 * GC_CHUNK * 32 + GC_CHUNK + GC_CHUNK + 1 + 15)  // records, offM, val, bal, pre (+ pad to 16 B)
enum { GR_RC_POS = 0, GR_RC_N = 3, GR_RC_IN = 6, GR_RC_L = 9, GR_RC_NEE = 12, GR_RC_MAT = 15, GR_JAC = 16, GR_LOBES = 17, GR_X1 = 18, GR_N1 = 21, GR_GATTR = 24, GR_FLUM = 25, GR_M = 26 };
__device__ __forceinline__ f3 rec3(const float* rec, int w, int c) { return f3{rec[w * 32 + c], rec[(w + 1) * 32 + c], rec[(w + 2) * 32 + c]}; }
__device__ __forceinline__ void rec3_store(float* rec, int w, int c, f3 v) { rec[w * 32 + c] = v.x, rec[(w + 1) * 32 + c] = v.y, rec[(w + 2) * 32 + c] = v.z; }

__global__ void __launch_bounds__(VRT_GRIS_THREADS, VRT_GRIS_MIN_BLOCKS) k_gris_coop(const __grid_constant__ Params P, RestirBuffers RB, uint32_t frame, int upper_in_smem, int fixed_words) {
  extern __shared__ uint32_t smem[];
  float4* s_mats = reinterpret_cast<float4*>(smem);
  for (int i = threadIdx.x; i < 128 * MAT_ROW_F4; i += blockDim.x) s_mats[i] = P.mats[i];
  float* s_unorm = reinterpret_cast<float*>(smem + 128 * MAT_ROW_F4 * 4);
  for (int i = threadIdx.x; i < 256; i += blockDim.x) s_unorm[i] = xdiv((float)i, 255.0f);
  const uint32_t* upper = P.upper;
  if (upper_in_smem) {
    uint32_t* s_upper = smem + fixed_words;
    for (int i = threadIdx.x; i < P.upper_words; i += blockDim.x) s_upper[i] = P.upper[i];
    upper = s_upper;
  }
  __syncthreads();

  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= P.n_tiles) return;  // warp-uniform: the whole warp leaves, the cooperative loops below always see 32 lanes
  const unsigned FULL = 0xffffffffu;
  uint32_t* ws = smem + ((fixed_words + (upper_in_smem ? P.upper_words : 0) + 3) & ~3) + (threadIdx.x >> 5) * GC_WARP_WORDS;
  float* rec = reinterpret_cast<float*>(ws);
  uint32_t* offM = ws + GC_REC_WORDS * 32;                              // [tap][centre]: (ox + 32) | (oy + 32) << 8 | f16 bits of the tap's M << 16
  float* val = reinterpret_cast<float*>(offM + GC_CHUNK * 32);          // [2][tap][centre]
  uint32_t* s_bal = reinterpret_cast<uint32_t*>(val + 2 * GC_CHUNK * 32);  // [tap] centres that accepted the tap
  uint32_t* s_pre = s_bal + GC_CHUNK;                                   // [tap + 1] pairs before the tap

  const int tile = P.tile_rank + P.tile_n * warp;
  const int u0 = (tile % P.tiles_x) * 8, v0 = (tile / P.tiles_x) * 4;
  const int u = u0 + (lane & 7), v = v0 + (lane >> 3);
  const int W = P.W, H = P.H;
  const size_t pidx = (size_t)v * W + u;

  GrisCtx G;
  G.mats = s_mats, G.unorm8 = s_unorm;
  G.cam_pos = P.cam_pos, G.light_dir = P.light_dir, G.sun_rad = P.light_weight * P.light_color;
  G.light_cos_max = P.light_cos_max, G.light_pdf_axis = cone_sample_pdf(P.light_cos_max, 1.0f);

  const float max_radius = 24.0f;
  const int max_taps = 32;
  const uint32_t key = path_key((uint32_t)pidx, frame, P.seed);
  const float4 gp = RB.gpos[pidx];
  const bool live = gp.w == 0.0f;
  const uint2 ga = RB.gattr[pidx];
  const uint32_t seed = hash3((uint32_t)u >> 3, (uint32_t)v >> 3, frame * 2u);
  const float angle_shift = (float)((seed & 0x007FFFFFu) | 0x3F800000u) / 4294967295.0f * VRT_PI;
  const float radius_shift = rnd(key, 65);
  const f3 center_x1{gp.x, gp.y, gp.z};
  const float center_dist = length(center_x1 - P.cam_pos);
  const f3 center_n1 = decode_unit_vector_3x16(h16val(ga.x), h16val(ga.x >> 16));
  float center_F_lum, center_M;
  {
    RReservoir center;
    load_reservoir(RB.reservoirs, pidx, s_unorm, center);
    center_F_lum = luminance(center.z.F), center_M = center.M;
    rec3_store(rec, GR_RC_POS, lane, center.z.rc_pos), rec3_store(rec, GR_RC_N, lane, center.z.rc_normal);
    rec3_store(rec, GR_RC_IN, lane, center.z.rc_incident_dir), rec3_store(rec, GR_RC_L, lane, center.z.rc_incident_L);
    rec3_store(rec, GR_RC_NEE, lane, center.z.rc_NEE_dir);
    rec[GR_RC_MAT * 32 + lane] = __uint_as_float(center.z.rc_mat_info);
    rec[GR_JAC * 32 + lane] = center.z.cached_jacobian_term;
    rec[GR_LOBES * 32 + lane] = __int_as_float(center.z.lobes);
    rec3_store(rec, GR_X1, lane, center_x1), rec3_store(rec, GR_N1, lane, center_n1);
    rec[GR_GATTR * 32 + lane] = __uint_as_float(ga.y);
    rec[GR_FLUM * 32 + lane] = center_F_lum;
    rec[GR_M * 32 + lane] = center_M;
  }
  float out_M = 0.0f, out_weight = 0.0f;
  int valid_samples = 0, sel_ti = -1;
  float canonical_mis_weight = 1.0f;

#pragma unroll 1
  for (int chunk = 0; chunk <= max_taps / GC_CHUNK; chunk++) {
    const bool resel = chunk == max_taps / GC_CHUNK;  // the extra pass-B round: one pair per lane, its selected tap
    uint32_t total = 32u;
    if (!resel) {
      // ---- accept: the similarity test of every (own pixel, tap) pair of this chunk
      total = 0u;
      __syncwarp();
#pragma unroll 1
      for (int j = 0; j < GC_CHUNK; j++) {
        const int i = chunk * GC_CHUNK + j;
        bool acc = false;
        uint32_t packed = 0u;
        do {
          if (!live) break;
          const float golden_angle = 2.399963229728f;
          const float angle = ((float)i + angle_shift) * golden_angle;
          const float offset_radius = sqrtf(((float)i + radius_shift) / (float)max_taps) * max_radius;
          float sn, cs;
          sincosf(angle, &sn, &cs);
          const int ox = (int)(cs * offset_radius), oy = (int)(sn * offset_radius);
          if (ox == 0 && oy == 0) break;
          const int tu = u + ox, tv = v + oy;
          if (tu < 0 || tv < 0 || tu >= W || tv >= H) break;
          const size_t ti = (size_t)tv * W + tu;
          const float4 ngp = __ldg(RB.gpos + ti);
          const uint2 nga = __ldg(RB.gattr + ti);
          if (ngp.w != 0.0f) break;
          const f3 neighbour_n1 = decode_unit_vector_3x16(h16val(nga.x), h16val(nga.x >> 16));
          const f3 neighbour_x1{ngp.x, ngp.y, ngp.z};
          const float neighbour_dist = length(neighbour_x1 - P.cam_pos);
          if (fabsf(neighbour_dist - center_dist) > 0.1f * center_dist || dot(center_n1, neighbour_n1) < 0.5f) break;
          acc = true;
          packed = (uint32_t)(ox + 32) | ((uint32_t)(oy + 32) << 8);
        } while (0);
        const uint32_t bal = __ballot_sync(FULL, acc);
        if (acc) offM[j * 32 + lane] = packed;
        if (lane == 0) s_bal[j] = bal, s_pre[j] = total;
        total += (uint32_t)__popc(bal);
      }
      if (lane == 0) s_pre[GC_CHUNK] = total;
      __syncwarp();

      // ---- pass A: the centre's sample shifted to every accepted tap
#pragma unroll 1
      for (uint32_t base = 0; base < total; base += 32u) {
        const uint32_t e = base + (uint32_t)lane;
        if (e < total) {
          int j = 0;
#pragma unroll
          for (int st = GC_CHUNK / 2; st > 0; st >>= 1)
            if (s_pre[j + st] <= e) j += st;
          const int c = (int)__fns(s_bal[j], 0u, (int)(e - s_pre[j]) + 1);
          const uint32_t pk = offM[j * 32 + c];
          const size_t ti = (size_t)(v0 + (c >> 3) + (int)((pk >> 8) & 255u) - 32) * W + (size_t)(u0 + (c & 7) + (int)(pk & 255u) - 32);
          const float4 ngp = __ldg(RB.gpos + ti);
          const uint2 nga = __ldg(RB.gattr + ti);
          const f3 neighbour_n1 = decode_unit_vector_3x16(h16val(nga.x), h16val(nga.x >> 16));
          const f3 neighbour_x1{ngp.x, ngp.y, ngp.z};
          int neighbour_mat_id;
          const Mat neighbour_mat = decode_material(G, nga.y, neighbour_mat_id);
          RReservoir cen;
          cen.M = 0.0f, cen.weight = 0.0f, cen.z.F = mk3(0.0f);
          cen.z.rc_pos = rec3(rec, GR_RC_POS, c), cen.z.rc_normal = rec3(rec, GR_RC_N, c), cen.z.rc_incident_dir = rec3(rec, GR_RC_IN, c);
          cen.z.rc_incident_L = rec3(rec, GR_RC_L, c), cen.z.rc_NEE_dir = rec3(rec, GR_RC_NEE, c);
          cen.z.rc_mat_info = __float_as_uint(rec[GR_RC_MAT * 32 + c]);
          cen.z.cached_jacobian_term = rec[GR_JAC * 32 + c];
          cen.z.lobes = __float_as_int(rec[GR_LOBES * 32 + c]);
          const size_t cpidx = (size_t)(v0 + (c >> 3)) * W + (size_t)(u0 + (c & 7));
          f3 c_d, c_s;
          float c_jacobian;
          shift_sample(G, neighbour_x1, neighbour_n1, neighbour_mat, prep_dst(G, neighbour_x1, neighbour_n1), cen, prep_rc(G, cen.z, RB.rc_skyT + cpidx), c_d, c_s,
                       c_jacobian);
          val[j * 32 + c] = luminance(c_d + c_s) * c_jacobian;
        }
      }
      __syncwarp();
    }

    // ---- pass B: every accepted tap's sample shifted to its centre (resel: the selected tap of every lane's own pixel)
#pragma unroll 1
    for (uint32_t base = 0; base < total; base += 32u) {
      const uint32_t e = base + (uint32_t)lane;
      int j = 0, c = lane;
      size_t ti = 0;
      bool work;
      if (resel) {
        work = sel_ti >= 0;
        ti = (size_t)(work ? sel_ti : 0);
      } else {
        work = e < total;
        if (work) {
#pragma unroll
          for (int st = GC_CHUNK / 2; st > 0; st >>= 1)
            if (s_pre[j + st] <= e) j += st;
          c = (int)__fns(s_bal[j], 0u, (int)(e - s_pre[j]) + 1);
          const uint32_t pk = offM[j * 32 + c];
          ti = (size_t)(v0 + (c >> 3) + (int)((pk >> 8) & 255u) - 32) * W + (size_t)(u0 + (c & 7) + (int)(pk & 255u) - 32);
        }
      }
      if (work) {
        RReservoir nb;
        load_reservoir(RB.reservoirs, ti, s_unorm, nb);
        const f3 cx1 = rec3(rec, GR_X1, c), cn1 = rec3(rec, GR_N1, c);
        int cmat_id;
        const Mat cmat = decode_material(G, __float_as_uint(rec[GR_GATTR * 32 + c]), cmat_id);
        f3 s_d, s_s;
        float jacobian;
        shift_sample(G, cx1, cn1, cmat, prep_dst(G, cx1, cn1), nb, prep_rc(G, nb.z, RB.rc_skyT + ti), s_d, s_s, jacobian);
        if (resel) {
          // what the in-order merge keeps of the selected tap (out.z = nb.z, out.z.F = s_d + s_s and the split integrand); parked in
          // the lane's own column of `val` (free now) so that nothing of it is live in registers across the loops
          rec3_store(val, 0, lane, nb.z.rc_pos), rec3_store(val, 3, lane, nb.z.rc_normal);
          rec3_store(val, 6, lane, s_d), rec3_store(val, 9, lane, s_s);
        } else {
          const float cM = rec[GR_M * 32 + c], cFlum = rec[GR_FLUM * 32 + c];
          const float center_p_hat = val[j * 32 + c];
          float canonical_weight = center_p_hat * nb.M;
          canonical_weight = canonical_weight / (center_p_hat * nb.M + cFlum * cM / (float)max_taps);
          const float p_hat = luminance(s_d + s_s);
          const float p_hat_from_neighbour = (RB.temporal ? luminance(nb.z.F) : p_hat) / jacobian;
          float neighbour_mis_weight = p_hat_from_neighbour * nb.M;
          neighbour_mis_weight = neighbour_mis_weight / (p_hat_from_neighbour * nb.M + p_hat * cM / (float)max_taps);
          if (isbad(neighbour_mis_weight)) neighbour_mis_weight = 0.0f;
          val[j * 32 + c] = nb.weight * p_hat * jacobian * neighbour_mis_weight;
          val[(GC_CHUNK + j) * 32 + c] = 1.0f - canonical_weight;
          offM[j * 32 + c] = (offM[j * 32 + c] & 0xffffu) | (h16bits(nb.M) << 16);
        }
      }
    }
    if (resel) break;
    __syncwarp();

    // ---- merge of this chunk's taps into the lane's own reservoir, in tap order (reservoir.py:76-86)
#pragma unroll 1
    for (int j = 0; j < GC_CHUNK; j++) {
      if (!((s_bal[j] >> lane) & 1u)) continue;
      const int i = chunk * GC_CHUNK + j;
      const uint32_t pk = offM[j * 32 + lane];
      const float in_w = val[j * 32 + lane];
      canonical_mis_weight += val[(GC_CHUNK + j) * 32 + lane];
      out_M += h16val(pk >> 16);
      if (in_w > 0.0f) {
        out_weight += in_w;
        if (rnd(key, 66u + (uint32_t)i) * out_weight <= in_w) sel_ti = (v + (int)((pk >> 8) & 255u) - 32) * W + (u + (int)(pk & 255u) - 32);
      }
      valid_samples += 1;
    }
  }

  f3 chosen_F_d = mk3(0.0f), chosen_F_s = mk3(0.0f), sel_rc_pos = mk3(0.0f), sel_rc_normal = mk3(0.0f);
  if (sel_ti >= 0) sel_rc_pos = rec3(val, 0, lane), sel_rc_normal = rec3(val, 3, lane), chosen_F_d = rec3(val, 6, lane), chosen_F_s = rec3(val, 9, lane);
  const f3 sel_F = chosen_F_d + chosen_F_s;
  f3 out_d, out_s = mk3(0.0f);
  RReservoir center;
  load_reservoir(RB.reservoirs, pidx, s_unorm, center);
  out_d = center.z.F;
  if (live) {
    int center_mat_id;
    const Mat center_mat = decode_material(G, ga.y, center_mat_id);
    // visibility of the resampled reconnection (pathtracer.py:957-965)
    bool force_add_canonical = false;
    if (out_weight > 0.0f) {
      const bool esc = is_vec_zero(sel_rc_normal);
      const f3 dir = esc ? sel_rc_pos : normalize(sel_rc_pos - center_x1);
      const f3 org = center_x1 + center_n1 * (0.003f * center_dist);
      Hit sh = next_hit<false>(P, upper, s_unorm, org, dir, true, nullptr, nullptr);
      const float actual_dist = esc ? VRT_INF : length(center_x1 - sel_rc_pos);
      if (sh.closest < VRT_INF && fabsf(sh.closest - actual_dist) > 0.1f * actual_dist) {
        out_weight = 0.0f;
        force_add_canonical = true;
      }
    }
    f3 out_F = sel_F;
    {
      const float in_w = center.weight * center_F_lum * canonical_mis_weight;
      out_M += center.M;
      if (in_w > 0.0f) {
        out_weight += in_w;
        if (rnd(key, 98) * out_weight <= in_w || force_add_canonical) {
          out_F = center.z.F;
          const float4 cd = RB.col_d[pidx], cs4 = RB.col_s[pidx];
          chosen_F_d = f3{cd.x, cd.y, cd.z};
          chosen_F_s = f3{cs4.x, cs4.y, cs4.z};
        }
      }
    }
    const float p_hat = luminance(out_F);  // finalize_without_M, then / (valid + 1)
    float Wt = p_hat < 1e-6f ? 0.0f : out_weight / p_hat;
    Wt = Wt / (float)(valid_samples + 1);
    const f3 emission = center_mat_id == 2 ? center_mat.base_col : mk3(0.0f);
    const float Wc = clampf(Wt, 0.0f, 50.0f);
    out_d = chosen_F_d * Wc + emission;
    out_s = chosen_F_s * Wc;
  }
  if (bad3(out_d)) out_d = mk3(0.0f);
  if (bad3(out_s)) out_s = mk3(0.0f);
  float4 a = P.accum[pidx];
  a.x += out_d.x + out_s.x, a.y += out_d.y + out_s.y, a.z += out_d.z + out_s.z, a.w += 1.0f;
  P.accum[pidx] = a;
}


// ---- launcher fragment (inside vrt_launch_gris, after `blocks` is known)
#if 0
  static const int coop = [] {
    const char* e = getenv("VRT_GRIS_COOP");
    return e ? atoi(e) : VRT_GRIS_COOP_DEFAULT;
  }();
  if (coop) {
    const size_t sm_coop = (((size_t)fixed_words + (uis ? (size_t)P.upper_words : 0) + 3) & ~(size_t)3) * 4 + (size_t)(VRT_GRIS_THREADS / 32) * GC_WARP_WORDS * 4;
    if (cudaError_t e = cudaFuncSetAttribute(k_gris_coop, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_coop)) return e;
    k_gris_coop<<<blocks, VRT_GRIS_THREADS, sm_coop, st>>>(P, RB, frame, uis, fixed_words);
  } else {
    k_gris<<<blocks, VRT_GRIS_THREADS, sm, st>>>(P, RB, frame, uis, fixed_words);
  }
#endif
