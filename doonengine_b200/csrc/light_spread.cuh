/* light_spread.cuh (included at the end of light.cu) -- the lighting update for SMALL dispatches: a voxel that traces specular rays gets
 * a whole warp, its rays spread over the lanes ("spread" lighting kernel); the other voxels (two short rays each) are lit one per lane
 * by their request's warp, exactly as in the warp-per-request kernel.
 *
 * A shader invocation traces up to 15 specular rays of up to specularBounceLimit segments, then per diffuse sample a path of up to
 * diffuseBounceLimit segments and a shadow ray -- one after the other (LI:244-261).  With one thread per voxel that is a serial chain
 * of ~35 dependent traversals, and a dispatch of a few hundred requests (the reference's own demo map: 782) cannot fill a B200 with
 * such chains: it ran at 8 % occupancy, as slow as the 16 CPU cores beside it (1.4 ms).  None of those rays depends on another one's
 * RESULT -- only on the order in which their contributions are added -- so here the rays of a voxel run side by side:
 *
 *     lane  0-14          specular ray i                               (LI:244-250)
 *     lane 15 + s         diffuse path of sample s, s < numDiffuseSamples  (LI:259)
 *     lane 15 + n + s     shadow ray of sample s                        (LI:260)          (n <= 8: at most 31 lanes)
 *
 * Every lane keeps its ray's contributions APART (AccList: at most specularBounceLimit addends for a specular ray, one for a diffuse
 * path, one for a shadow ray) and lane 0 then replays them into the two running sums in exactly the shader's order: specular ray 0's
 * addends, ray 1's, ...; ambient, diffuse, shadow of sample 0, of sample 1, ... -- the same additions on the same values, so the staged
 * words are bit-identical to the other kernels' (tests/test_parity_gpu.py runs every lighting test against this kernel too).
 *
 * What DOES carry from ray to ray inside a shader invocation is the "inside a transparent block" state (lastVoxID / lastVoxRefract,
 * SH:324-325) and the guard flag; a ray that ends inside glass changes how the next one starts.  Every lane first starts clean, as the
 * first ray does; then a lane whose predecessor (in the shader's order) ended inside glass traces its ray again from that state, until
 * no lane's start differs from its predecessor's end.  On the demo map, where glass panes touch floors and walls, the earlier scheme --
 * discard everything and let lane 0 light the voxel serially -- made those voxels the tail of the whole dispatch.
 */
#define SPREAD_WARPS 4
/* resident CTAs per SM the register allocation aims for (A/B builds: 5 = 96 registers, 6 = 80, 8 = 64) */
#ifndef SPREAD_MIN_BLOCKS
#define SPREAD_MIN_BLOCKS 5
#endif

template <int DUMMY>
__global__ void __launch_bounds__(SPREAD_WARPS * 32, SPREAD_MIN_BLOCKS) dn_light_spread_kernel(DnbScene S, const uint32_t* __restrict__ requests, DnbWork W, DnbStagingTargets T)
{
	__shared__ float s_add[SPREAD_WARPS][32][DNB_MAX_ADDENDS][3];
	__shared__ uint32_t s_num[SPREAD_WARPS][32];
	__shared__ DnbSlot s_slot[SPREAD_WARPS];
	const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	const uint32_t numRequests = work_requests(W);
	/* work items, one per warp: first every request of this launch (its voxels WITHOUT specular rays, one per lane, the plain way:
	 * two short rays each -- and the empty rows of the request), then every voxel of every request (its rays spread over the lanes;
	 * taken only if it traces specular rays).  Both kinds decide with the same predicate, so every staged row has exactly one writer. */
	const uint32_t requestItems = work_ctas(W, numRequests) * 4u;
	const uint32_t totalItems = requestItems * 33u;
	const uint32_t n = c_light.numDiffuseSamples;

	for(uint32_t item = blockIdx.x * SPREAD_WARPS + warp; item < totalItems; item += gridDim.x * SPREAD_WARPS)
	{
		if(item < requestItems)
		{
			const uint32_t r = (W.firstCta + (item >> 2) * W.ctaStride) * 4u + (item & 3u);
			if(r < numRequests)
				light_request<false, true>(S, requests, r, warp, lane, T, s_slot);
			__syncwarp();
			continue;
		}
		const uint32_t j = item - requestItems; /* voxel j & 31 of request (j >> 5) */

		/* ---- set-up, LI:207-232: every lane computes the same (the loads are broadcasts) ---- */
		const uint32_t r = (W.firstCta + (j >> 7) * W.ctaStride) * 4u + ((j >> 5) & 3u);
		if(r >= numRequests)
			continue;
		const uint32_t request = __ldg(requests + r);
		const uint32_t mapIndex = request >> 4;
		const size_t at = (size_t)r * 96u + (j & 31u);
		const uint32_t slotId = __ldg(S.tileSlot + mapIndex) - 1u;
		const DnbSlot* slot = S.slots + slotId;
		const uint32_t voxNum = (j & 31u) + (request & 15u) * 32u;
		const int local = slotId == 0xFFFFFFFFu ? -1 : flat_nth_voxel(slot, voxNum);
		if(local < 0)
			continue; /* (the request's own work item has staged the empty row) */
		const uint4 rec = __ldg(S.records + (__ldg(&slot->voxelBase) + voxNum));
		const f3 normal = vox_normal(rec);
		const f3 albedo = vox_albedo(rec);
		const DnbMaterial material = load_material(S, vox_material(rec));
		const uint32_t ns = __ldg(&slot->numSamples);
		const float indirectSamples = (float)(ns < c_light.maxDiffuseSamples ? ns : c_light.maxDiffuseSamples);
		const i3 chunkPos = {local & 7, (local >> 3) & 7, local >> 6};
		const i3 mapPos = {__ldg(&slot->pos[0]), __ldg(&slot->pos[1]), __ldg(&slot->pos[2])};
		f3 rayPos = (tof3(chunkPos) * 0.125f + tof3(mapPos)) + 0.0625f;
		rayPos = rayPos + normal * (0.0625f - DNB_EPSILON);
		if(!traces_specular(material, rayPos, normal))
			continue; /* lit by its request's work item */

		LightCtx cx;
		ray_state_reset(cx.st);
		cx.lc = DnbCounters{0, 0, 0, 0, 0, 0, 0};
		cx.firstSample = indirectSamples == 0.0f;
		cx.sourceVisible = (__ldg(S.visible + (mapIndex >> 5)) >> (mapIndex & 31u)) & 1u;

		const f3 viewDir = rayPos - ld3(c_light.camPos);
		const bool diffuse = material.specular < 1.0f;

		/* ---- this lane's ray ----
		 * Every ray is first traced as if it started outside any transparent block (lastVoxID = 255), which is how the first ray starts.
		 * Then each lane looks at the state its PREDECESSOR in the shader's order ended in (specular 0..14, then diffuse path and shadow
		 * ray of sample 0, of sample 1, ...); a lane whose predecessor ended inside glass traces its ray again from that state, which
		 * may change how IT ends, and so on until nothing changes -- at most one round per ray, normally none or one. */
		const bool hasRay = lane < 15u || (diffuse && lane < 15u + 2u * n);
		const uint32_t predLane = lane == 0u ? 0u : lane < 15u ? lane - 1u : lane == 15u ? 14u : lane < 15u + n ? lane + n - 1u : lane - n;
		uint32_t startID = 255u;
		float startRefract = 1.0f;
		bool redo = hasRay, broken = false;
		AccList acc;
		acc.n = 0;
		acc.numTiles = 0;
#pragma unroll 1
		for(uint32_t round = 0; round < 32u; round++)
		{
			if(redo)
			{
				ray_state_reset(cx.st);
				cx.st.lastVoxID = startID;
				cx.st.lastVoxRefract = startRefract;
				acc.n = 0;
				acc.numTiles = 0;
				if(lane < 15u)
				{
					const f3 reflected = reflect3(normalize3(viewDir), normal);
					const f3 specDir = normalize3(reflected * (float)material.shininess + ld3(c_spherePoints[lane])) + DNB_EPSILON;
					specular_ray<false>(S, cx, rayPos, specDir, albedo, material.reflectType, acc);
				}
				else if(lane < 15u + n)
					diffuse_ray<false>(S, cx, normal + DNB_EPSILON, rayPos, rec, lane - 15u, acc);
				else
					shadow_ray<false>(S, cx, rayPos, lane - 15u - n, acc);
				broken = broken || cx.st.tripped || acc.n > (uint32_t)DNB_MAX_ADDENDS || acc.numTiles > (uint32_t)DNB_MAX_ADDENDS;
			}
			const uint32_t predID = __shfl_sync(0xFFFFFFFFu, cx.st.lastVoxID, predLane);
			const float predRefract = __shfl_sync(0xFFFFFFFFu, cx.st.lastVoxRefract, predLane);
			redo = hasRay && lane != 0u && (predID != startID || __float_as_uint(predRefract) != __float_as_uint(startRefract));
			if(redo)
			{
				startID = predID;
				startRefract = predRefract;
			}
			if(!__any_sync(0xFFFFFFFFu, redo) || __any_sync(0xFFFFFFFFu, broken))
				break;
		}

		/* a guard that tripped stays tripped for the rest of the invocation, and a ray with more addends than a lane keeps cannot be
		 * replayed: lane 0 lights such a voxel the serial way (never seen on the test maps) */
		if(__any_sync(0xFFFFFFFFu, broken || redo))
		{
			if(lane == 0)
			{
				ray_state_reset(cx.st);
				uint32_t w1, w2, w3;
				light_voxel<false>(S, cx, rec, material, rayPos, indirectSamples, w1, w2, w3);
				stage_words(T, at, w1, w2, w3);
			}
			continue;
		}

		/* the rays stand: their visible-bit propagations (LI:101-105) take effect */
		{
			AccSum sink;
#pragma unroll
			for(uint32_t k = 0; k < DNB_MAX_ADDENDS; k++)
				if(k < acc.numTiles)
					sink.hit_tile(S, acc.tiles[k]);
		}
#pragma unroll
		for(uint32_t k = 0; k < DNB_MAX_ADDENDS; k++)
		{
			s_add[warp][lane][k][0] = acc.a[k].x;
			s_add[warp][lane][k][1] = acc.a[k].y;
			s_add[warp][lane][k][2] = acc.a[k].z;
		}
		s_num[warp][lane] = acc.n;
		__syncwarp();

		/* ---- replay in the shader's order (LI:244-263) ---- */
		if(lane == 0)
		{
			f3 specLight = splat3(0.0f), diffuseLight = splat3(0.0f);
			{
				for(uint32_t i = 0; i < 15u; i++)
					for(uint32_t k = 0; k < s_num[warp][i]; k++)
						specLight = specLight + mk3(s_add[warp][i][k][0], s_add[warp][i][k][1], s_add[warp][i][k][2]);
				specLight = div3(specLight, 15.0f);
			}
			if(diffuse)
			{
				const f3 ambient = ld3(S.ambient);
				for(uint32_t s = 0; s < n; s++)
				{
					diffuseLight = diffuseLight + ambient;
					if(s_num[warp][15u + s])
						diffuseLight = diffuseLight + mk3(s_add[warp][15u + s][0][0], s_add[warp][15u + s][0][1], s_add[warp][15u + s][0][2]);
					if(s_num[warp][15u + n + s])
						diffuseLight = diffuseLight + mk3(s_add[warp][15u + n + s][0][0], s_add[warp][15u + n + s][0][1], s_add[warp][15u + n + s][0][2]);
				}
				diffuseLight = div3(vox_diffuse(rec) * indirectSamples + diffuseLight, indirectSamples + (float)n);
			}
			uint32_t w1, w2, w3;
			pack_lit_words(rec, specLight, diffuseLight, w1, w2, w3);
			stage_words(T, at, w1, w2, w3);
		}
		__syncwarp(); /* the addend buffers are rewritten by the next voxel */
	}
}

/* usable when the rays of a voxel fit a warp and no specular ray can produce more addends than a lane keeps */
extern "C" bool dnb_light_spread_usable(uint32_t numDiffuseSamples, uint32_t specularBounceLimit)
{
	return 15u + 2u * numDiffuseSamples <= 32u && specularBounceLimit <= (uint32_t)DNB_MAX_ADDENDS;
}

extern "C" cudaError_t dnb_launch_light_spread(const DnbScene* scene, const uint32_t* requests, const DnbWork* work, uint32_t gridCtas, const DnbStagingTargets* targets, cudaStream_t stream)
{
	if(gridCtas == 0)
		return cudaSuccess;
	/* gridCtas counts 4-request CTAs = 4 request items + 128 voxel items, a warp each; the kernel strides, so the grid is only capped */
	const unsigned long long want = (unsigned long long)gridCtas * (132u / SPREAD_WARPS);
	const uint32_t grid = (uint32_t)(want < 148ull * 64ull ? want : 148ull * 64ull);
	{ DNB_LAUNCHED(1); dn_light_spread_kernel<0><<<grid, SPREAD_WARPS * 32, 0, stream>>>(*scene, requests, *work, *targets); }
	return cudaGetLastError();
}
