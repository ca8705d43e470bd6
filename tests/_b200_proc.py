"""The reference-side binding of INTEGRATION.md, shared by the CPU (oracle
test double) and GPU (real engine) drop-in tests: subclasses of the UNCHANGED
reference procedures that override only the `sampling` cached_property."""
import functools


def make_dmc_proc(dmc_exec, b200_dmc, **sampling_kwargs):
    import attr

    @attr.s(auto_attribs=True, frozen=True)
    class B200DMCProc(dmc_exec.Proc):

        @functools.cached_property
        def sampling(self):
            nts = self.num_time_steps_block
            den = ssf = None
            if self.should_eval_density:
                den = b200_dmc.DensityEstSpec(self.density_spec.num_bins,
                                              self.density_spec.as_pure_est,
                                              nts)
            if self.should_eval_ssf:
                ssf = b200_dmc.SSFEstSpec(self.ssf_spec.num_modes,
                                          self.ssf_spec.as_pure_est, nts)
            return b200_dmc.Sampling(
                self.model_spec, self.time_step, self.max_num_walkers,
                self.target_num_walkers, self.num_walkers_control_factor,
                self.rng_seed, density_est_spec=den, ssf_est_spec=ssf,
                **sampling_kwargs)

    return B200DMCProc


def make_vmc_proc(vmc_exec, b200_vmc, **sampling_kwargs):
    import attr

    @attr.s(auto_attribs=True, frozen=True)
    class B200VMCProc(vmc_exec.Proc):

        @functools.cached_property
        def sampling(self):
            ssf = None
            if self.should_eval_ssf:
                ssf = b200_vmc.SSFEstSpec(self.ssf_spec.num_modes)
            return b200_vmc.Sampling(self.model_spec, self.move_spread,
                                     self.rng_seed, ssf_est_spec=ssf,
                                     **sampling_kwargs)

    return B200VMCProc
