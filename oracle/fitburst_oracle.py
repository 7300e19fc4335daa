"""
TEST INFRASTRUCTURE ONLY -- the CPU oracle for the fitburst hot path.

Nothing under fitburst_b200/ may import this module. It is used by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs, and only as
the checker / the CPU baseline, never as the thing shipped.

This is a channel-vectorised numpy restatement of the reference's algorithm (CHIMEFRB/fitburst,
pure Python). Each function cites the reference file:line it follows ("ref:" = path under
/root/reference/fitburst). It is pinned against the UNMODIFIED reference run in the build
container (tests/test_oracle_vs_reference.py, oracle/make_golden.py -> tests/golden/*.npz);
the reference itself ships no golden vectors or known-answer tests for this path
(SURVEY.md section 4), so those frozen outputs ARE the pin. Generated with numpy 2.3.5 /
scipy 1.18.1.

Third-party arithmetic on the path (not vendored in the reference, pyproject.toml:16-23 gives
only lower bounds): scipy.special.erfcx/erfc (ref: routines/profile.py:107,112),
scipy.optimize.least_squares defaults = TRF without bounds (ref: analysis/fitter.py:300),
numpy.linalg.inv/solve (ref: analysis/fitter.py:482,484, routines/ism.py:46).
"""
import numpy as np
import scipy.special as ss
from scipy.optimize import least_squares

DISPERSION_CONSTANT = 4149.377593360996  # ref: backend/general.yaml:7

PARAMETER_ORDER = [  # ref: analysis/model.py:92-102
    "amplitude",
    "arrival_time",
    "burst_width",
    "dm",
    "dm_index",
    "scattering_timescale",
    "scattering_index",
    "spectral_index",
    "spectral_running",
]
FITTER_GLOBAL_PARAMETERS = ["dm", "dm_index", "scattering_timescale", "scattering_index"]  # ref: fitter.py:72


def upsample_1d(array_orig, diff_orig, factor):
    """ref: routines/manipulate.py:75-107."""
    diff_new = diff_orig / factor
    bound_lo = array_orig[0] - (diff_orig / 2) + (diff_new / 2)
    bound_hi = array_orig[-1] + (diff_orig / 2) - (diff_new / 2)
    return np.linspace(bound_lo, bound_hi, num=len(array_orig) * factor)


def downsample_last(array, factor):
    """ref: routines/manipulate.py:10-41, applied along the last axis."""
    shape = array.shape[:-1] + (array.shape[-1] // factor, factor)
    return np.sum(array.reshape(shape), axis=-1) / factor


def dm_delay(dm_value, dm_const, dm_idx, freq1, freq2):
    """ref: routines/ism.py:50-80."""
    return dm_value * dm_const * (freq1 ** dm_idx - freq2 ** dm_idx)


def spectrum_rpl(freqs, freq_ref, sp_idx, sp_run):
    """ref: routines/spectrum.py:11-43."""
    log_freq = np.log(freqs / freq_ref)
    return np.exp(sp_idx * log_freq + sp_run * log_freq ** 2)


def profile_gaussian(values, mean, width):
    """ref: routines/profile.py:15-49 (normalize=False, the only in-tree use)."""
    return np.exp(-0.5 * ((values - mean) / width) ** 2)


def profile_pbf(time, width, freq, ref_freq, sc_time_ref, sc_index, normalize):
    """
    ref: routines/profile.py:52-114 with toa = 0 (model.py:221). `time` has shape
    (..., Kf, N) and `freq` (..., Kf, 1). The erfcx/erfc regime is chosen per time column:
    erfcx only where arg > -20 for every sub-frequency (profile.py:103, 147).
    """
    sc_time = sc_time_ref * (freq / ref_freq) ** sc_index
    z = time / width
    ratio = width / sc_time
    arg1 = (ratio - z) / np.sqrt(2)
    amp_term = np.sqrt(np.pi / 2) * ratio if normalize else sc_time_ref / sc_time
    use_erfcx = np.all(arg1 > -20.0, axis=-2, keepdims=True)
    with np.errstate(over="ignore", invalid="ignore", under="ignore"):
        regime1 = amp_term * np.exp(-0.5 * z ** 2) * ss.erfcx(arg1)
        regime2 = amp_term * np.exp((ratio / 2 - z) * ratio) * ss.erfc(arg1)
    return np.where(use_erfcx, regime1, regime2)


class OracleModel:
    """
    Restatement of SpectrumModeler (ref: analysis/model.py:15-409). Holds the same four
    (Nf, Nt, Nc) caches; compute_model() is vectorised over channels in chunks instead of the
    reference's python loop over (channel, component) at model.py:147-150.
    """

    def __init__(self, freqs, times, dm_incoherent=0.0, factor_freq_upsample=1,
                 factor_time_upsample=1, num_components=1, is_dedispersed=False, is_folded=False,
                 scintillation=False, normalize_pbf=False, chunk=64):
        self.freqs = np.asarray(freqs, dtype=float)
        self.times = np.asarray(times, dtype=float)
        self.dm_incoherent = dm_incoherent
        self.factor_freq_upsample = factor_freq_upsample
        self.factor_time_upsample = factor_time_upsample
        self.num_components = num_components
        self.is_dedispersed = is_dedispersed
        self.is_folded = is_folded
        self.scintillation = scintillation
        self.normalize_pbf = normalize_pbf  # ref: backend/general.yaml:12, read at model.py:337
        self.chunk = chunk
        self.num_freq = len(self.freqs)
        self.num_time = len(self.times)
        self.res_freq = np.fabs(self.freqs[1] - self.freqs[0])  # model.py:87
        self.res_time = np.fabs(self.times[1] - self.times[0])  # model.py:88
        self.parameters = list(PARAMETER_ORDER)
        for name in self.parameters:
            setattr(self, name, None)
        shape = (self.num_freq, self.num_time, self.num_components)
        self.amplitude_per_component = np.zeros(shape)
        self.spectrum_per_component = np.zeros(shape)
        self.timediff_per_component = np.zeros(shape)
        self.timeprof_per_component = np.zeros(shape)

    # ref: analysis/model.py:383-409
    def update_parameters(self, model_parameters, global_parameters=("dm", "scattering_timescale")):
        for name, values in model_parameters.items():
            setattr(self, name, values)
            if len(values) > 1:
                self.num_components = len(values)
                if name in global_parameters:
                    setattr(self, name, [values[0]] * self.num_components)

    # ref: analysis/model.py:354-381
    def get_parameters_dict(self):
        out = {name: getattr(self, name) for name in self.parameters}
        out["ref_freq"] = self.ref_freq
        return out

    def _subfreqs(self, freqs):
        """ref: model.py:177-181 + manipulate.py:99-105, one linspace row per channel."""
        kf = self.factor_freq_upsample
        diff_new = self.res_freq / kf
        lo = freqs - (self.res_freq / 2) + (diff_new / 2)
        hi = freqs + (self.res_freq / 2) - (diff_new / 2)
        return np.linspace(lo, hi, num=kf, axis=-1)  # (chunk, Kf)

    def _profile(self, times, sc_time_ref, sc_index, width, freqs, ref_freq):
        """
        ref: model.py:281-352. `times` (chunk, Kf, N), `freqs` (chunk, Kf, 1). The reference
        decides "scattered or Gaussian" per (channel, component) call (model.py:339).
        """
        times_copy = times
        if self.is_folded:  # model.py:327-332
            num = times.shape[-1]
            res_time = times[..., 1] - times[..., 0]
            start = times[..., -1] + res_time
            stop = times[..., -1] + res_time * num
            extended = np.linspace(start, stop, num=num, axis=-1)
            times_copy = np.concatenate([times, extended], axis=-1)
        with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
            sc_time = sc_time_ref * (freqs / ref_freq) ** sc_index
        scattered = np.any(sc_time > 0.0, axis=(-2, -1))  # (chunk,)
        profile = np.empty(times_copy.shape)
        if np.any(scattered):
            sel = times_copy[scattered]
            sel = np.where(sel < -5 * width, -5 * width, sel)  # model.py:340
            profile[scattered] = profile_pbf(
                sel, width, freqs[scattered], ref_freq, sc_time_ref, sc_index, self.normalize_pbf
            )
        if np.any(~scattered):
            profile[~scattered] = profile_gaussian(times_copy[~scattered], 0.0, width)  # model.py:345
        if self.is_folded:  # model.py:349-350
            num = times.shape[-1]
            profile = profile.reshape(profile.shape[:-1] + (2, num)).mean(-2)
        return profile

    def compute_model(self, data=None):
        """ref: analysis/model.py:126-279."""
        if self.scintillation and data is None:
            raise SystemExit("ERROR: scintillation modelling is desired by data are missing!")
        kt = self.factor_time_upsample
        for lo in range(0, self.num_freq, self.chunk):
            hi = min(lo + self.chunk, self.num_freq)
            freqs = self.freqs[lo:hi]
            freq_arr = self._subfreqs(freqs)  # (chunk, Kf)
            for comp in range(self.num_components):
                amplitude = self.amplitude[comp]
                arrival_time = self.arrival_time[comp]
                dm = self.dm[0]
                dm_index = self.dm_index[0]
                ref_freq = self.ref_freq[comp]
                sc_idx = self.scattering_index[0]
                sc_time = self.scattering_timescale[0]
                sp_idx = self.spectral_index[comp]
                sp_run = self.spectral_running[comp]
                width = self.burst_width[comp]
                # model.py:184-199
                delay = dm_delay(self.dm_incoherent + dm, DISPERSION_CONSTANT, dm_index, freq_arr, ref_freq)
                delay -= dm_delay(self.dm_incoherent, DISPERSION_CONSTANT, dm_index, freqs, ref_freq)[:, None]
                # model.py:202-209
                current_times = upsample_1d(self.times - arrival_time, self.res_time, kt)
                times_arr = current_times[None, None, :] - delay[:, :, None]  # (chunk, Kf, Nt*Kt)
                # model.py:212-216
                self.timediff_per_component[lo:hi, :, comp] = downsample_last(times_arr.mean(axis=1), kt)
                # model.py:219-233
                profile = self._profile(times_arr, sc_time, sc_idx, width, freq_arr[:, :, None], ref_freq)
                self.timeprof_per_component[lo:hi, :, comp] = downsample_last(profile.mean(axis=1), kt)
                # model.py:236-256
                profile = profile * spectrum_rpl(freq_arr, ref_freq, sp_idx, sp_run)[:, :, None]
                profile = downsample_last(profile.mean(axis=1), kt)
                self.amplitude_per_component[lo:hi, :, comp] = (
                    spectrum_rpl(freqs, ref_freq, sp_idx, sp_run) * (10 ** amplitude)
                )[:, None]
                self.spectrum_per_component[lo:hi, :, comp] = (10 ** amplitude) * profile
        if self.scintillation:  # model.py:267-277 + ism.py:11-48
            for freq in range(self.num_freq):
                prof = self.timeprof_per_component[freq]  # (Nt, Nc)
                gram = np.zeros((self.num_components, self.num_components))
                rhs = np.zeros(self.num_components)
                for c1 in range(self.num_components):
                    rhs[c1] = np.sum(data[freq] * prof[:, c1])
                    for c2 in range(self.num_components):
                        gram[c1, c2] = np.sum(prof[:, c1] * prof[:, c2])
                amps = np.linalg.solve(gram, rhs)
                self.amplitude_per_component[freq, :, :] = amps[None, :]
                self.spectrum_per_component[freq, :, :] = amps[None, :] * prof
        return np.sum(self.spectrum_per_component, axis=2)


# --------------------------------------------------------------------------------------------
# first derivatives, ref: routines/derivative.py:14-936
# --------------------------------------------------------------------------------------------

def _argument_erf(width, timediff, sc_time):
    """ref: derivative.py:14-38."""
    return (timediff - width ** 2 / sc_time) / width / np.sqrt(2)


def _argument_exp(width, timediff, sc_time):
    """ref: derivative.py:40-69 (evaluated as written, not as -T^2/2w^2)."""
    arg_erf = _argument_erf(width, timediff, sc_time)
    return 0.5 * (width / sc_time) ** 2 - timediff / sc_time - arg_erf ** 2


def _amplitude_pbf(freq, parameters, component, normalize):
    """ref: derivative.py:71-103."""
    ref_freq = parameters["ref_freq"][component]
    amp = (freq / ref_freq) ** -parameters["scattering_index"][0]
    if normalize:
        amp = amp * np.sqrt(np.pi / 2.0) * parameters["burst_width"][component] / parameters["scattering_timescale"][0]
    return amp


def _deriv_amplitude_pbf(name, freq, parameters, component, normalize):
    """ref: derivative.py:105-152."""
    width = parameters["burst_width"][component]
    ref_freq = parameters["ref_freq"][component]
    sc_index = parameters["scattering_index"][0]
    sc_time = parameters["scattering_timescale"][0]
    out = 0.0
    if normalize:
        sc_time_freq = sc_time * (freq / ref_freq) ** sc_index
        if name == "burst_width":
            out = np.sqrt(np.pi / 2) / sc_time_freq
        elif name == "scattering_index":
            out = -np.sqrt(np.pi / 2) * np.log(freq / ref_freq) * width / sc_time_freq
        elif name == "scattering_timescale":
            out = -np.sqrt(np.pi / 2) * width / (sc_time_freq * sc_time)
    elif name == "scattering_index":
        out = -np.log(freq / ref_freq) * (freq / ref_freq) ** -sc_index
    return out


def _deriv_time_dm(name, freq, ref_freq, dm, dm_index):
    """ref: derivative.py:211-250. Note: `dm` is the fitted DM only, not dm_incoherent + dm."""
    if name == "dm":
        return -DISPERSION_CONSTANT * (freq ** dm_index - ref_freq ** dm_index)
    if name == "dm_index":
        return -DISPERSION_CONSTANT * dm * (np.log(freq) * freq ** dm_index - np.log(ref_freq) * ref_freq ** dm_index)
    return 0.0


def _deriv_argument_erf(name, freq, timediff, parameters, component):
    """ref: derivative.py:252-307."""
    width = parameters["burst_width"][component]
    dm = parameters["dm"][0]
    dm_index = parameters["dm_index"][0]
    ref_freq = parameters["ref_freq"][component]
    sc_time = parameters["scattering_timescale"][0]
    sc_index = parameters["scattering_index"][0]
    sc_time_freq = sc_time * (freq / ref_freq) ** sc_index
    if name == "arrival_time":
        return -1 / width / np.sqrt(2)
    if name == "burst_width":
        return -(timediff / width ** 2 + 1 / sc_time_freq) / np.sqrt(2)
    if name in ("dm", "dm_index"):
        return _deriv_time_dm(name, freq, ref_freq, dm, dm_index) / width / np.sqrt(2)
    if name == "scattering_timescale":
        return width / sc_time / sc_time_freq / np.sqrt(2)
    if name == "scattering_index":
        return np.log(freq / ref_freq) * width / sc_time_freq / np.sqrt(2)
    return 0.0


class OracleDerivatives:
    """
    First derivatives of the model, ref: routines/derivative.py:454-936. The reference loops over
    channels in four of the nine functions; the formulas are element-wise so they are evaluated
    here on whole (Nf, Nt) maps, with the per-channel `sc_times_freq == 0.0` branch as a mask.
    """

    def __init__(self, model: OracleModel):
        self.model = model
        self.normalize = model.normalize_pbf

    def _scattered_term3(self, parameters, comp, name, amp_component):
        model = self.model
        width = parameters["burst_width"][comp]
        ref_freq = parameters["ref_freq"][comp]
        sc_times_freq = parameters["scattering_timescale"][0] * (model.freqs / ref_freq) ** parameters["scattering_index"][0]
        timediff = model.timediff_per_component[:, :, comp]
        amp_pbf = _amplitude_pbf(model.freqs, parameters, comp, self.normalize)[:, None]
        with np.errstate(all="ignore"):
            arg_exp = _argument_exp(width, timediff, sc_times_freq[:, None])
            deriv_arg_erf = _deriv_argument_erf(name, model.freqs[:, None], timediff, parameters, comp)
            term3 = model.amplitude_per_component[:, :, comp] * amp_pbf * np.exp(arg_exp) * deriv_arg_erf * 2 / np.sqrt(np.pi)
            # reference quirk: the four global functions pass the OUTER component here (derivative.py:707,779,847,916)
            deriv_amp_pbf = _deriv_amplitude_pbf(name, model.freqs[:, None], parameters, amp_component, self.normalize)
            term1 = model.spectrum_per_component[:, :, comp] * deriv_amp_pbf / amp_pbf
        return term1, term3, sc_times_freq, timediff

    def amplitude(self, parameters, component=0):
        return np.log(10) * self.model.spectrum_per_component[:, :, component]  # :454-476

    def spectral_running(self, parameters, component=0):
        log_freq = np.log(self.model.freqs / parameters["ref_freq"][component]) ** 2  # :479-503
        return log_freq[:, None] * self.model.spectrum_per_component[:, :, component]

    def spectral_index(self, parameters, component=0):
        log_freq = np.log(self.model.freqs / parameters["ref_freq"][component])  # :505-529
        return log_freq[:, None] * self.model.spectrum_per_component[:, :, component]

    def burst_width(self, parameters, component=0):
        """:531-590."""
        width = parameters["burst_width"][component]
        current_model = self.model.spectrum_per_component[:, :, component]
        term1, term3, sc_tf, timediff = self._scattered_term3(parameters, component, "burst_width", component)
        with np.errstate(all="ignore"):
            scattered = term1 + (width / sc_tf[:, None] ** 2) * current_model + term3
        gaussian = timediff ** 2 * current_model / width ** 3
        return np.where((sc_tf == 0.0)[:, None], gaussian, scattered)

    def arrival_time(self, parameters, component=0):
        """:592-652."""
        width = parameters["burst_width"][component]
        current_model = self.model.spectrum_per_component[:, :, component]
        term1, term3, sc_tf, timediff = self._scattered_term3(parameters, component, "arrival_time", component)
        with np.errstate(all="ignore"):
            scattered = term1 + current_model / sc_tf[:, None] + term3
        gaussian = timediff * current_model / width ** 2
        return np.where((sc_tf == 0.0)[:, None], gaussian, scattered)

    def _dm_like(self, name, parameters, component, add_all):
        """:654-725 (dm) and :727-798 (dm_index)."""
        model = self.model
        out = np.zeros(model.spectrum_per_component.shape)
        for comp in range(model.timediff_per_component.shape[2]):
            width = parameters["burst_width"][comp]
            current_model = model.spectrum_per_component[:, :, comp]
            ref_freq = parameters["ref_freq"][comp]
            deriv_td = _deriv_time_dm(name, model.freqs, ref_freq, parameters["dm"][0], parameters["dm_index"][0])
            term1, term3, sc_tf, timediff = self._scattered_term3(parameters, comp, name, component)
            with np.errstate(all="ignore"):
                scattered = term1 + (-deriv_td[:, None] * current_model / sc_tf[:, None]) + term3
            gaussian = -timediff * current_model / width ** 2 * deriv_td[:, None]
            out[:, :, comp] = np.where((sc_tf == 0.0)[:, None], gaussian, scattered)
        return np.sum(out, axis=2) if add_all else out[:, :, component]

    def dm(self, parameters, component=0, add_all=True):
        return self._dm_like("dm", parameters, component, add_all)

    def dm_index(self, parameters, component=0, add_all=True):
        return self._dm_like("dm_index", parameters, component, add_all)

    def scattering_timescale(self, parameters, component=0, add_all=True):
        """:800-867 (no Gaussian branch: sc_time == 0 divides by zero in the reference)."""
        model = self.model
        out = np.zeros(model.spectrum_per_component.shape)
        sc_time = parameters["scattering_timescale"][0]
        for comp in range(model.timediff_per_component.shape[2]):
            width = parameters["burst_width"][comp]
            current_model = model.spectrum_per_component[:, :, comp]
            term1, term3, sc_tf, timediff = self._scattered_term3(parameters, comp, "scattering_timescale", component)
            with np.errstate(all="ignore"):
                term2 = (-(width / sc_tf[:, None]) ** 2 + timediff / sc_tf[:, None]) * current_model / sc_time
            out[:, :, comp] = term1 + term2 + term3
        return np.sum(out, axis=2) if add_all else out[:, :, component]

    def scattering_index(self, parameters, component=0, add_all=True):
        """:869-936."""
        model = self.model
        out = np.zeros(model.spectrum_per_component.shape)
        for comp in range(model.timediff_per_component.shape[2]):
            width = parameters["burst_width"][comp]
            current_model = model.spectrum_per_component[:, :, comp]
            log_freq = np.log(model.freqs / parameters["ref_freq"][comp])
            term1, term3, sc_tf, timediff = self._scattered_term3(parameters, comp, "scattering_index", component)
            with np.errstate(all="ignore"):
                term2 = -log_freq[:, None] * current_model * ((width / sc_tf[:, None]) ** 2 - timediff / sc_tf[:, None])
            out[:, :, comp] = term1 + term2 + term3
        return np.sum(out, axis=2) if add_all else out[:, :, component]

    def first(self, name, parameters, component=0, add_all=True):
        if name in FITTER_GLOBAL_PARAMETERS:
            return getattr(self, name)(parameters, component, add_all)
        return getattr(self, name)(parameters, component)


# --------------------------------------------------------------------------------------------
# fitter, ref: analysis/fitter.py
# --------------------------------------------------------------------------------------------

class OracleFitter:
    """Restatement of LSFitter (ref: analysis/fitter.py:16-542)."""

    def __init__(self, data, model: OracleModel, good_freq, weighted_fit=True, weight_range=None):
        self.data = data
        self.model = model
        self.good_freq = np.asarray(good_freq, dtype=bool)
        self.fit_statistics = {}
        if model.scintillation:  # fitter.py:60-65
            self.fit_parameters = [p for p in model.parameters if p not in ("amplitude", "spectral_index", "spectral_running")]
        else:
            self.fit_parameters = list(model.parameters)
        self.global_parameters = list(FITTER_GLOBAL_PARAMETERS)
        self.success = None
        self.weighted_fit = weighted_fit
        self.weight_range = weight_range
        self.deriv = OracleDerivatives(model)
        self.hessian_fn = None  # set by oracle.hessian_oracle when available
        self._set_weights()

    def _set_weights(self):
        """ref: fitter.py:506-542. The default range [0, Nt-1] is used as a slice: last sample excluded."""
        idx_range = self.weight_range
        if idx_range is None:
            idx_range = [0, len(self.data[0, :]) - 1]
        with np.errstate(all="ignore"):
            std = np.sqrt(np.mean(self.data[:, idx_range[0]: idx_range[1]] ** 2, axis=1))
            self.weights = np.empty_like(std)
            self.weights[self.good_freq] = 1.0 / std[self.good_freq] if self.weighted_fit else 1.0
        self.weights[~self.good_freq] = 0.0

    def fix_parameter(self, parameter_list):
        for name in parameter_list:  # fitter.py:348-349
            self.fit_parameters.remove(name)

    def get_fit_parameters_list(self):
        """ref: fitter.py:353-384."""
        out = []
        for name in self.model.parameters:
            if name in self.fit_parameters:
                sub = getattr(self.model, name)
                out += [sub[0]] if name in self.global_parameters else list(sub)
        return out

    def load_fit_parameters_list(self, parameter_list):
        """ref: fitter.py:386-430."""
        out, idx = {}, 0
        for name in self.model.parameters:
            if name in self.fit_parameters:
                if name in self.global_parameters:
                    out[name] = [parameter_list[idx]]
                    idx += 1
                else:
                    out[name] = parameter_list[idx: idx + self.model.num_components]
                    idx += self.model.num_components
        return out

    def compute_residuals(self, parameter_list):
        """ref: fitter.py:231-266."""
        self.model.update_parameters(self.load_fit_parameters_list(parameter_list))
        model = self.model.compute_model(data=self.data)
        return ((self.data - model) * self.weights[:, None]).ravel()

    def compute_jacobian(self, parameter_list):
        """ref: fitter.py:180-229 (differentiates at the model's CURRENT state, :202)."""
        parameters = self.model.get_parameters_dict()
        jac = np.zeros((self.model.num_freq * self.model.num_time, len(parameter_list)))
        col = 0
        for name in self.fit_parameters:
            num = 1 if name in self.global_parameters else self.model.num_components
            for comp in range(num):
                deriv = -self.deriv.first(name, parameters, comp)
                jac[:, col] = (deriv * self.weights[:, None]).ravel()
                col += 1
        return jac

    def fit(self, exact_jacobian=True, with_statistics=True):
        """ref: fitter.py:268-322 (scipy defaults: TRF, no bounds, ftol=xtol=gtol=1e-8)."""
        x0 = self.get_fit_parameters_list()
        self.results = least_squares(self.compute_residuals, x0, jac=self.compute_jacobian if exact_jacobian else "2-point")
        if with_statistics:
            self._compute_fit_statistics(self.results)
        return self.results

    def _compute_fit_statistics(self, fit_result):
        """ref: fitter.py:432-504."""
        stats = self.fit_statistics
        num_fit = len(fit_result.x)
        stats["num_freq"] = self.model.num_freq
        stats["num_freq_good"] = int(np.sum(self.good_freq))
        stats["num_fit_parameters"] = num_fit
        stats["num_observations"] = stats["num_freq_good"] * int(self.model.num_time) - num_fit
        stats["num_time"] = self.model.num_time
        chisq_initial = float(np.sum((self.data * self.weights[:, None]) ** 2))
        chisq_final = float(np.sum(fit_result.fun ** 2))
        stats["chisq_initial"] = chisq_initial
        stats["chisq_final"] = chisq_final
        stats["chisq_final_reduced"] = chisq_final / stats["num_observations"]
        stats["snr"] = float(np.sqrt(chisq_initial - chisq_final))
        stats["bestfit_parameters"] = self.load_fit_parameters_list(fit_result.x.tolist())
        stats["bestfit_uncertainties"] = None
        stats["bestfit_covariance"] = None
        self.hessian_approx = fit_result.jac.T.dot(fit_result.jac)
        self.covariance_approx = np.linalg.inv(self.hessian_approx)
        if self.hessian_fn is not None:
            self.hessian, self.covariance_labels = self.hessian_fn(self, self.data)
            self.covariance = np.linalg.inv(0.5 * self.hessian)
            with np.errstate(invalid="ignore"):
                uncertainties = [float(x) for x in np.sqrt(np.diag(self.covariance)).tolist()]
            if np.any(np.isnan(uncertainties)):
                uncertainties = [float(x) for x in np.sqrt(np.diag(self.covariance_approx)).tolist()]
            stats["bestfit_uncertainties"] = self.load_fit_parameters_list(uncertainties)
