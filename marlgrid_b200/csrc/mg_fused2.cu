// mg_fused2.cu -- dispatch of the specialised fused step+observe kernel (mg_fused2.cuh; instantiated per observation mode
// and view size in mg_fused2_{enc,rgb}{7,5}.cu).
#include "mg_common.cuh"

namespace mg {

template <int OBS, int V>
int launch_fused2_ov(const KP& p, cudaStream_t s);
extern template int launch_fused2_ov<1, 7>(const KP&, cudaStream_t);
extern template int launch_fused2_ov<1, 5>(const KP&, cudaStream_t);
extern template int launch_fused2_ov<2, 7>(const KP&, cudaStream_t);
extern template int launch_fused2_ov<2, 5>(const KP&, cudaStream_t);

template <int V>
int launch_fused2_hide(const KP& p, cudaStream_t s);
extern template int launch_fused2_hide<7>(const KP&, cudaStream_t);
extern template int launch_fused2_hide<5>(const KP&, cudaStream_t);

template <int V>
int launch_fused2_ks(const KP& p, int n_steps, cudaStream_t s);
extern template int launch_fused2_ks<7>(const KP&, int, cudaStream_t);
extern template int launch_fused2_ks<5>(const KP&, int, cudaStream_t);

// the specialised kernel's own reset routes (and the background generator) sample agents over the whole grid in a standard world
static bool standard_worlds(const KP& p) { return p.scenario == 0 && p.ax0 == 0 && p.ay0 == 0 && p.aw == p.W && p.ah == p.H && p.amax == 100000; }
static bool al16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; }

int launch_fused2(const KP& p, int obs, cudaStream_t s) {
  if (!fused2_eligible(p) || p.A > 6 || (p.V != 7 && p.V != 5)) return MG_E_UNSUPPORTED;
  if (p.hide != 0u && obs != 1) return MG_E_UNSUPPORTED;  // hide_item_types with RGB observations: general kernels
  if (!standard_worlds(p)) return MG_E_UNSUPPORTED;  // spawn boxes / other generators: the general kernels' sequential reset
  if (!al16(p.actions) || !al16(p.rewards) || !al16(p.done) || !al16(p.obs)) return MG_E_UNSUPPORTED;  // bulk copies need 16-byte alignment
  if (obs == 1 && p.hide != 0u) return p.V == 7 ? launch_fused2_hide<7>(p, s) : launch_fused2_hide<5>(p, s);
  if (obs == 1) return p.V == 7 ? launch_fused2_ov<1, 7>(p, s) : launch_fused2_ov<1, 5>(p, s);
  // RGB: tile size 8 (every registered env), rotation-equivariant atlas (one slot per tile), tile ids that fit a byte
  if (obs == 2 && p.ts == 8 && p.orient_slots == 1 && p.n_tiles < 255 && al16(p.atlas))
    return p.V == 7 ? launch_fused2_ov<2, 7>(p, s) : launch_fused2_ov<2, 5>(p, s);
  return MG_E_UNSUPPORTED;
}

// n_steps env.steps in ONE launch on per-step slices actions[t] / rewards[t] / done[t] / obs[t] (encoded observations); the
// state of every tile stays in shared memory between the steps.  MG_E_UNSUPPORTED: shape or batch size outside the
// persistent kernel's reach (the caller then launches step by step).
int launch_fused2_rollout(const KP& p, int n_steps, cudaStream_t s) {
  if (n_steps < 1 || !fused_eligible(p) || p.hide != 0u || (p.V != 7 && p.V != 5) || !standard_worlds(p)) return MG_E_UNSUPPORTED;
  if (!al16(p.actions) || !al16(p.rewards) || !al16(p.done) || !al16(p.obs)) return MG_E_UNSUPPORTED;
  if (((long long)p.B * p.A * 4) % 16 != 0 || p.B % 16 != 0) return MG_E_UNSUPPORTED;  // per-step slices must stay 16-byte aligned
  return p.V == 7 ? launch_fused2_ks<7>(p, n_steps, s) : launch_fused2_ks<5>(p, n_steps, s);
}

}  // namespace mg
