/* peer.cu -- what the replicas of a sharded volume do to each other's memory besides the stores issued by the
 * lighting and draw kernels themselves (light.cu, draw.cu): a device-side barrier and the visible-bitmap merge.
 *
 * The reference has no multi-GPU path; the design is SURVEY.md 8e with the NCCL all-gathers
 * replaced by loads and stores on peer memory mapped over NVLink / NVSwitch (DoonEngine/b200.h "multi-GPU over
 * peer memory").
 *
 * Barrier: every replica owns a mailbox of DNB_MAX_PEERS epoch words.  Replica r arrives by storing the epoch
 * into slot r of EVERY replica's mailbox (st.release.sys, after a system-scope fence: everything the preceding
 * kernels of this stream stored into peer memory is visible to whoever sees the epoch) and leaves once all slots of
 * its OWN mailbox have reached the epoch (ld.acquire.sys on local memory: the spin costs no NVLink traffic).
 * Epochs only grow, so a mailbox never needs resetting, and a replica that is a whole phase ahead cannot be
 * mistaken for one that has arrived (the comparison is on the signed difference).  The spin is bounded: after
 * DNB_BARRIER_TIMEOUT_NS the kernel gives up and raises status[1], so a dead peer cannot hang the GPU.
 */
#include "kernels.h"

#define DNB_BARRIER_TIMEOUT_NS 20000000000ull /* 20 s */

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v)
{
	asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p)
{
	uint32_t v;
	asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}

__device__ __forceinline__ unsigned long long global_timer_ns()
{
	unsigned long long t;
	asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
	return t;
}

/* one CTA of DNB_MAX_PEERS threads; thread p talks to replica p.  status[0] = last epoch completed, status[1] = time-outs */
__global__ void __launch_bounds__(32) dn_peer_barrier_kernel(DnbPeerTable T, uint32_t epoch, uint32_t* __restrict__ status)
{
	const uint32_t p = threadIdx.x;
	bool ok = true;
	if(p < T.world)
	{
		__threadfence_system();
		st_release_sys(T.mailbox[p] + T.rank, epoch);

		const uint32_t* mine = T.mailbox[T.rank] + p;
		const unsigned long long t0 = global_timer_ns();
		uint32_t spins = 0;
		while((int32_t)(ld_acquire_sys(mine) - epoch) < 0)
		{
			if((++spins & 1023u) == 0 && global_timer_ns() - t0 > DNB_BARRIER_TIMEOUT_NS)
			{
				ok = false;
				break;
			}
			__nanosleep(64);
		}
	}
	const uint32_t allOk = __all_sync(0xFFFFFFFFu, ok);
	if(p == 0)
	{
		if(allOk)
			status[0] = epoch;
		else
			atomicAdd(status + 1, 1u);
	}
}

/* visible[i] |= every other replica's visible[i].  The peers' words are read with volatile (uncached) loads: peer
 * addresses bypass the local L2, and a stale L1 line must not satisfy them. */
__global__ void __launch_bounds__(256) dn_peer_or_visible_kernel(DnbPeerTable T, uint32_t* __restrict__ visible, uint32_t words)
{
	const uint32_t i = blockIdx.x * 256 + threadIdx.x;
	if(i >= words)
		return;
	uint32_t v = visible[i], add = 0;
	for(uint32_t p = 0; p < T.world; p++)
		if(p != T.rank)
			add |= *reinterpret_cast<const volatile uint32_t*>(T.visible[p] + i);
	if(add & ~v)
		visible[i] = v | add;
}

extern "C" cudaError_t dnb_launch_peer_barrier(const DnbPeerTable* peers, uint32_t epoch, uint32_t* status, cudaStream_t stream)
{
	{ DNB_LAUNCHED(1); dn_peer_barrier_kernel<<<1, 32, 0, stream>>>(*peers, epoch, status); }
	return cudaGetLastError();
}

extern "C" cudaError_t dnb_launch_peer_or_visible(const DnbPeerTable* peers, uint32_t* visible, uint32_t words, cudaStream_t stream)
{
	if(words == 0)
		return cudaSuccess;
	{ DNB_LAUNCHED(1); dn_peer_or_visible_kernel<<<(words + 255) / 256, 256, 0, stream>>>(*peers, visible, words); }
	return cudaGetLastError();
}
