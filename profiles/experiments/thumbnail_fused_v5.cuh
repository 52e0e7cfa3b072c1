/* thumbnail_fused_v5.cuh -- v4 (thumbnail_fused_mma.cuh) with the warp roles folded.
 *
 * What v4's profile shows: its V warps spend 28% of their time waiting for a TMA stage
 * although the ring depth does not matter (6 stages == 4 stages) and the same access
 * pattern feeds 7.3 TB/s when nothing is computed (tools/tma_feed_bench.cu).  The wait is
 * skew: a stage is refilled only once EVERY V warp has read it, so all V warps run at the
 * pace of the slowest, and the slowest are the ones whose SM sub-partition (warp slot % 4)
 * holds the most V warps.  6 V warps + 1 H warp + 1 producer warp cannot be spread evenly
 * over 4 sub-partitions, and the hardware rotates the first warp slot from CTA to CTA
 * (tools/warp_slots.cu), so no static role placement fixes it.
 *
 * v5 has one kind of warp.  Every warp
 *     consumes TMA stages (premultiply + box sums -> row quads),
 *     runs the reducev MMAs of its own 64 / CPT columns,
 *     and does 1 / NW of the reduceh + unpremultiply + store work of the PREVIOUS chunk;
 * warp 0 additionally re-arms a stage one step after it was released (1 instruction for an
 * interior stage: a tiled-TMA box).  With the warp count a multiple of 4 (CPT 1: 12 warps
 * for 384 columns) every sub-partition carries the same load whatever the rotation.
 *
 * Shared memory: stages | full[S] empty[S] shfull[2] shempty[2] | quadbuf | sh[2][8][..] | hcoef | uscale
 */

template <int VS, int NP, bool PREMUL, int HSQ, int WCOLS, int CPT>
__global__ void __launch_bounds__(WCOLS / CPT, WCOLS <= 448 ? 2 : 1)
thumbnail_fused_v5_kernel(const __grid_constant__ FusedParams P, const __grid_constant__ CUtensorMap tmap, int use_tmap,
	const uint8_t *__restrict__ in, size_t in_frame_stride, uint8_t *__restrict__ out, size_t out_frame_stride, int frame0)
{
	extern __shared__ __align__(128) unsigned char smem_raw[];

	constexpr int K = kV4Rows;
	constexpr int S = V4Stages<VS>::value;
	constexpr int PITCH = (WCOLS + 8) * 4;
	constexpr int NPR = NP > 0 ? NP : 1;
	constexpr int HSHIFT = HSQ == 2 ? 1 : HSQ == 4 ? 2 : 3;
	constexpr int rows_per_stage = 2 * VS;
	constexpr unsigned stage_bytes = (unsigned) rows_per_stage * PITCH;

	const int NT = (int) blockDim.x;
	const int NW = NT / 32;
	const int NC = NT * CPT; /* columns */
	const int t = (int) threadIdx.x;
	const int warp = t >> 5, lane = t & 31;
	const bool lane0 = lane == 0;
	const int NPh = NP > 0 ? NP : P.NPh;
	const int shs = NC / HSQ / 2; /* pairs per sh row */
	const unsigned QS = (unsigned) NC * 16u + 16u; /* bytes per quad slot */

	unsigned char *stages = smem_raw;
	uint64_t *bars = (uint64_t *) (smem_raw + S * stage_bytes);
	unsigned char *quadbuf = (unsigned char *) (bars + 2 * S + 4);
	uint2 *sh = (uint2 *) (quadbuf + (size_t) kV4Quads * QS);
	int *hcoef = (int *) (sh + (size_t) 2 * K * shs);
	int *uscale = hcoef + P.nhsets * P.NPh;

	const unsigned stages_s = smem_addr(stages);
	const unsigned full_s = smem_addr(bars);
	const unsigned empty_s = full_s + 8u * S;
	const unsigned shfull_s = empty_s + 8u * S;
	const unsigned shempty_s = shfull_s + 16u;

	if (t == 0) {
		for (int i = 0; i < S; i++) {
			mbar_init(full_s + 8u * i, 1);
			mbar_init(empty_s + 8u * i, NW);
		}
		for (int i = 0; i < 2; i++) {
			mbar_init(shfull_s + 8u * i, NW);
			mbar_init(shempty_s + 8u * i, NW);
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	for (int i = t; i < P.nhsets * P.NPh; i += NT)
		hcoef[i] = P.hcoef[i];
	if (PREMUL)
		for (int i = t; i < 256; i += NT)
			uscale[i] = i == 0 ? 0 : (int) __ddiv_rn(__dmul_rn(256.0, 255.0), (double) i);

	const int xa = blockIdx.x * P.TW;
	const int xb = min(xa + P.TW, P.OW);
	const int bw = xb - xa;
	const int y_begin = blockIdx.y * P.RPC;
	const int y_end = min(y_begin + P.RPC, P.OH);
	const int frame = frame0 + blockIdx.z;
	const uint8_t *fin = in + (size_t) frame * in_frame_stride;
	uint8_t *fout = out + (size_t) frame * out_frame_stride;

	const int pair_h0 = __ldg(&P.hcol[xa]).x;
	const int E0 = 2 * pair_h0 + P.hgrid;
	const int NE = 2 * (__ldg(&P.hcol[xb - 1]).x + P.NPh - pair_h0);

	auto column_of = [&](int tt) {
		const int e = E0 + tt / HSQ;
		const int k = tt - (tt / HSQ) * HSQ;
		const int sc = max(0, min(e - P.hembed, P.Ws - 1));
		return min(sc * HSQ + k, P.W - 1);
	};
	const int c_lo = column_of(0) & ~3;
	const int c_hi = min(P.W, (column_of(NE * HSQ - 1) + 4) & ~3);
	const unsigned row_bytes = (unsigned) (c_hi - c_lo) * 4u;
	const int q_first = __ldg(&P.vchunk[y_begin / K]).x;
	const int p_first = 2 * q_first;												  /* first stage (pair of shrunk rows) */
	const int p_last = 2 * __ldg(&P.vchunk[(y_end - 1) / K]).y + 1;					  /* last stage */
	const uint8_t *src0 = fin + (size_t) c_lo * 4;

	__syncthreads();

	/* warp 0: arm stage slot `slot` with pair p (its barrier must have been released) */
	auto arm_stage = [&](int p, int slot) {
		const int sr0 = 2 * p - P.vembed;
		const bool interior = use_tmap && sr0 >= 0 && sr0 + 1 <= P.Hs - 1 && (sr0 + 2) * VS <= P.H;
		if (interior) {
			/* 2 VS consecutive rows, none an edge replica: one tiled-TMA box {PITCH bytes, 2 VS rows} */
			if (lane0) {
				mbar_expect_tx(full_s + 8u * slot, stage_bytes);
				asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
								 stages_s + (unsigned) slot * stage_bytes),
							 "l"(&tmap), "r"(c_lo >> 1), "r"(sr0 * VS), "r"(frame), "r"(full_s + 8u * slot)
							 : "memory");
			}
		}
		else {
			if (lane0)
				mbar_expect_tx(full_s + 8u * slot, (unsigned) rows_per_stage * row_bytes);
			__syncwarp();
			if (lane < rows_per_stage) {
				const int j = lane / VS, k = lane - j * VS;
				const int sr = max(0, min(2 * p + j - P.vembed, P.Hs - 1));
				const int row = min(sr * VS + k, P.H - 1);
				bulk_copy_g2s(stages_s + (unsigned) slot * stage_bytes + (unsigned) lane * PITCH, src0 + (size_t) row * P.in_bpl,
					row_bytes, full_s + 8u * slot);
			}
		}
	};
	if (warp == 0)
		for (int i = 0; i < S && p_first + i <= p_last; i++)
			arm_stage(p_first + i, i);

	/* CPT 2: the thread's two columns are adjacent and start on an even column (one 64-bit LDS per row) */
	const unsigned char *my_cols = stages + (size_t) (column_of(min(t * CPT, NE * HSQ - CPT)) - c_lo) * 4u;
	const unsigned accm = P.accmul;
	unsigned k16;
	asm volatile("mov.u32 %0, 0x10000;" : "=r"(k16));
	int k20;
	asm volatile("mov.u32 %0, 0x100000;" : "=r"(k20));
	const unsigned amend2 = (unsigned) (VS / 2) * (VB200_V4_HADD2 ? 1u : accm) * 0x00010001u;
	/* MMA fragment coordinates */
	const int tig = lane & 3, g = lane >> 2, ch = g & 3, jj = g >> 2;
	const int warp_col0 = warp * 32 * CPT;
	const unsigned char *a_base = quadbuf + (size_t) tig * QS + (size_t) (warp_col0 + 4 * jj) * 16u + (unsigned) ch * 4u;
	unsigned char *q_store = quadbuf + (size_t) (t * CPT) * 16u;
	const int ch_off = ch == 0 ? 0 : ch == 1 ? 4 : ch == 2 ? 2 : 6; /* sh pair layout [rA rB bA bB gA gB aA aB] */
	const int warp_sx0 = warp_col0 / HSQ;

	/* reduceh + unpremultiply + store of chunk cc (rows ya ..): this warp's 1 / NW of the outputs */
	auto h_pass = [&](int cc, int ya) {
		const int rows = min(K, y_end - ya);
		const int buf = cc & 1;
		const uint2 *shb = sh + (size_t) buf * K * shs;
		mbar_wait(shfull_s + 8u * buf, (unsigned) (cc >> 1) & 1u);
		for (int idx = t; idx < rows * bw; idx += NT) {
			const int k = fast_div(idx, bw);
			const int x = xa + (idx - k * bw);
			const int2 hc = __ldg(&P.hcol[x]);
			const uint2 *win = shb + k * shs + (hc.x - pair_h0);
			const int *cfp = hcoef + hc.y * NPh;
			int r = VB200_INTERPOLATE_SCALE >> 1, gg = r, b = r, a = r;
			if (NP > 0) {
#pragma unroll
				for (int kk = 0; kk < NPR; kk++) {
					const uint2 w = win[kk];
					const unsigned c = (unsigned) cfp[kk];
					r = dp2a_lo(c, w.x, r);
					b = dp2a_hi(c, w.x, b);
					gg = dp2a_lo(c, w.y, gg);
					a = dp2a_hi(c, w.y, a);
				}
			}
			else
				for (int kk = 0; kk < NPh; kk++) {
					const uint2 w = win[kk];
					const unsigned c = (unsigned) cfp[kk];
					r = dp2a_lo(c, w.x, r);
					b = dp2a_hi(c, w.x, b);
					gg = dp2a_lo(c, w.y, gg);
					a = dp2a_hi(c, w.y, a);
				}
			r = max(0, min(r >> VB200_INTERPOLATE_SHIFT, 255));
			gg = max(0, min(gg >> VB200_INTERPOLATE_SHIFT, 255));
			b = max(0, min(b >> VB200_INTERPOLATE_SHIFT, 255));
			a = max(0, min(a >> VB200_INTERPOLATE_SHIFT, 255));
			if (PREMUL) {
				const int sc = uscale[a];
				r = ((r * sc + 128) >> 8) & 0xff;
				gg = ((gg * sc + 128) >> 8) & 0xff;
				b = ((b * sc + 128) >> 8) & 0xff;
			}
			*(unsigned *) (fout + (size_t) (ya + k) * P.out_bpl + (size_t) x * 4) =
				(unsigned) r | ((unsigned) gg << 8) | ((unsigned) b << 16) | ((unsigned) a << 24);
		}
		__syncwarp();
		if (lane0)
			mbar_arrive(shempty_s + 8u * buf);
	};

	int n = 0; /* stages consumed so far: slot n % S, use n / S */
	int s = 0;
	unsigned phase = 0;
	int qdone = q_first;
	int chunk = 0;

	for (int ya = y_begin; ya < y_end; ya += K, chunk++) {
		const int q1 = __ldg(&P.vchunk[ya / K]).y;

		for (int q = qdone; q <= q1; q++) {
			unsigned rb[4][CPT], ga[4][CPT];
#pragma unroll
			for (int r = 0; r < 4; r++)
#pragma unroll
				for (int i = 0; i < CPT; i++)
					rb[r][i] = ga[r][i] = amend2;
#pragma unroll
			for (int half = 0; half < 2; half++) {
				const unsigned soff = (unsigned) s * stage_bytes;
				mbar_wait(full_s + 8u * s, phase);
				if (CPT == 2) {
					uint2 pa[VS], pb[VS];
#pragma unroll
					for (int k = 0; k < VS; k++) {
						pa[k] = *(const uint2 *) (my_cols + soff + k * PITCH);
						pb[k] = *(const uint2 *) (my_cols + soff + (VS + k) * PITCH);
					}
					__syncwarp();
					if (lane0)
						mbar_arrive(empty_s + 8u * s);
#pragma unroll
					for (int k = 0; k < VS; k++) {
						V4_ACC(pa[k].x, rb[2 * half][0], ga[2 * half][0]);
						V4_ACC(pa[k].y, rb[2 * half][CPT - 1], ga[2 * half][CPT - 1]);
						V4_ACC(pb[k].x, rb[2 * half + 1][0], ga[2 * half + 1][0]);
						V4_ACC(pb[k].y, rb[2 * half + 1][CPT - 1], ga[2 * half + 1][CPT - 1]);
					}
				}
				else {
					unsigned pa[VS], pb[VS];
#pragma unroll
					for (int k = 0; k < VS; k++) {
						pa[k] = *(const unsigned *) (my_cols + soff + k * PITCH);
						pb[k] = *(const unsigned *) (my_cols + soff + (VS + k) * PITCH);
					}
					__syncwarp();
					if (lane0)
						mbar_arrive(empty_s + 8u * s);
#pragma unroll
					for (int k = 0; k < VS; k++) {
						V4_ACC(pa[k], rb[2 * half][0], ga[2 * half][0]);
						V4_ACC(pb[k], rb[2 * half + 1][0], ga[2 * half + 1][0]);
					}
				}
				/* warp 0 re-arms the slot released ONE step ago (every warp has read it by now, or is about to:
				 * the wait is short) with the stage S steps after it
				 */
				if (warp == 0 && n > 0) {
					const int np = n - 1;
					if (p_first + np + S <= p_last) {
						const int slot = s == 0 ? S - 1 : s - 1;
						mbar_wait(empty_s + 8u * slot, (unsigned) (np / S) & 1u);
						arm_stage(p_first + np + S, slot);
					}
				}
				n++;
				if (++s == S) {
					s = 0;
					phase ^= 1u;
				}
			}
			/* box averages are bytes 1 and 3 of each lane word: transpose 4 rows into quads */
#pragma unroll
			for (int i = 0; i < CPT; i++) {
				if (VB200_V4_HADD2) {
#pragma unroll
					for (int r = 0; r < 4; r++) {
						rb[r][i] *= accm;
						ga[r][i] *= accm;
					}
				}
				const unsigned rb01 = __byte_perm(rb[0][i], rb[1][i], 0x7351); /* [r0 r1 b0 b1] */
				const unsigned rb23 = __byte_perm(rb[2][i], rb[3][i], 0x7351);
				const unsigned ga01 = __byte_perm(ga[0][i], ga[1][i], 0x7351);
				const unsigned ga23 = __byte_perm(ga[2][i], ga[3][i], 0x7351);
				uint4 w;
				w.x = __byte_perm(rb01, rb23, 0x5410); /* r rows 0..3 */
				w.y = __byte_perm(ga01, ga23, 0x5410); /* g */
				w.z = __byte_perm(rb01, rb23, 0x7632); /* b */
				w.w = __byte_perm(ga01, ga23, 0x7632); /* a */
				*(uint4 *) (q_store + (size_t) (q & (kV4Quads - 1)) * QS + i * 16) = w;
			}
		}
		qdone = max(qdone, q1 + 1);
		__syncwarp();

		/* reducev on the tensor pipe + in-thread shrinkh into sh[buf] (free once every warp has done h_pass(chunk - 2)) */
		const int buf = chunk & 1;
		const uint4 bf = __ldg(&P.vbfrag[(size_t) (ya / K) * 32 + lane]); /* {hi b0, hi b1, lo b0, lo b1} */
		mbar_wait(shempty_s + 8u * buf, ((unsigned) (chunk >> 1) & 1u) ^ 1u);
		unsigned char *shb = (unsigned char *) (sh + (size_t) buf * K * shs) + (size_t) (2 * tig) * shs * 8 + ch_off;
#pragma unroll 2
		for (int tp = 0; tp < 4 * CPT; tp++) {
			unsigned a[2][4];
			int dh[2][4], dl[2][4];
#pragma unroll
			for (int T = 0; T < 2; T++) {
				const unsigned char *ap = a_base + tp * 128 + T * 32;
				a[T][0] = *(const unsigned *) (ap);
				a[T][1] = *(const unsigned *) (ap + 16);
				a[T][2] = *(const unsigned *) (ap + 4 * (size_t) QS);
				a[T][3] = *(const unsigned *) (ap + 4 * (size_t) QS + 16);
			}
#pragma unroll
			for (int T = 0; T < 2; T++) {
#pragma unroll
				for (int i = 0; i < 4; i++) {
					dh[T][i] = 0;
					dl[T][i] = VB200_INTERPOLATE_SCALE >> 1;
				}
				mma_u8s8(dh[T], a[T], bf.x, bf.y);
				mma_u8u8(dl[T], a[T], bf.z, bf.w);
			}
#pragma unroll
			for (int r = 0; r < 2; r++) {
				const int v00 = v4_finish(dh[0][r], dl[0][r], k20);
				const int v01 = v4_finish(dh[0][2 + r], dl[0][2 + r], k20);
				const int v10 = v4_finish(dh[1][r], dl[1][r], k20);
				const int v11 = v4_finish(dh[1][2 + r], dl[1][2 + r], k20);
				if (HSQ == 2) {
					const int sx = warp_sx0 + tp * 4 + 2 * jj;
					unsigned char *d = shb + (size_t) r * shs * 8 + (sx >> 1) * 8;
					d[0] = (unsigned char) ((v00 + v01 + 1) >> 1);
					d[1] = (unsigned char) ((v10 + v11 + 1) >> 1);
				}
				else if (HSQ == 4) {
					const int sx = warp_sx0 + tp * 2 + jj;
					shb[(size_t) r * shs * 8 + (sx >> 1) * 8 + (sx & 1)] = (unsigned char) ((v00 + v01 + v10 + v11 + 2) >> 2);
				}
				else {
					int sum = v00 + v01 + v10 + v11;
					sum += __shfl_xor_sync(0xffffffffu, sum, 16);
					const int sx = warp_sx0 + tp;
					if (jj == 0)
						shb[(size_t) r * shs * 8 + (sx >> 1) * 8 + (sx & 1)] = (unsigned char) ((sum + 4) >> HSHIFT);
				}
			}
		}
		__syncwarp();
		if (lane0)
			mbar_arrive(shfull_s + 8u * buf);

		if (chunk > 0)
			h_pass(chunk - 1, ya - K);
	}
	h_pass(chunk - 1, y_begin + (chunk - 1) * K);
}
