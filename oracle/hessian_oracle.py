"""
TEST INFRASTRUCTURE ONLY -- CPU oracle for the exact chi^2 Hessian of the fitburst model.

Restates the 45 mixed partial derivatives of routines/derivative.py:938-2836 (reference paths
are relative to /root/reference/fitburst) as whole-map numpy expressions, and
LSFitter.compute_hessian (analysis/fitter.py:81-178) on top of them. The reference's quirks are
reproduced on purpose (they define bestfit_uncertainties): `ref_freq ** 2` in dm_dm_index
(derivative.py:2267-2268), Gaussian-form expressions and the component-SUMMED dm_index first
derivative in arrival_time_dm_index / dm_dm_index (:2082-2115, :2236-2273), amp_pbf of the outer
component in dm_index_dm_index (:2807), sums over components for the global x global pairs.
Pinned against the unmodified reference by tests/test_oracle_vs_reference.py.
"""
import numpy as np

from .fitburst_oracle import (DISPERSION_CONSTANT, FITTER_GLOBAL_PARAMETERS, OracleDerivatives,
                              _amplitude_pbf, _argument_erf, _argument_exp, _deriv_amplitude_pbf,
                              _deriv_argument_erf, _deriv_time_dm)

SQ2 = np.sqrt(2)


def _deriv_argument_exp(name, freq, timediff, parameters, component):
    """ref: derivative.py:309-371."""
    width = parameters["burst_width"][component]
    dm = parameters["dm"][0]
    dm_index = parameters["dm_index"][0]
    ref_freq = parameters["ref_freq"][component]
    sc_time = parameters["scattering_timescale"][0]
    sc_index = parameters["scattering_index"][0]
    sc_time_freq = sc_time * (freq / ref_freq) ** sc_index
    arg_erf = _argument_erf(width, timediff, sc_time_freq)
    deriv_arg_erf = _deriv_argument_erf(name, freq, timediff, parameters, component)
    out = 0.0
    if name == "arrival_time":
        out = 1 / sc_time_freq
    elif name == "burst_width":
        out = width / sc_time_freq ** 2
    elif name in ("dm", "dm_index"):
        out = -_deriv_time_dm(name, freq, ref_freq, dm, dm_index) / sc_time_freq
    elif name == "scattering_timescale":
        out = -(width / sc_time_freq) ** 2 / sc_time + timediff / sc_time_freq / sc_time
    elif name == "scattering_index":
        out = np.log(freq / ref_freq) * (-(width / sc_time_freq) ** 2 + timediff / sc_time_freq)
    return out - 2 * arg_erf * deriv_arg_erf


def _deriv2_argument_erf(name1, name2, freq, timediff, parameters, component):
    """ref: derivative.py:373-452 (only the 12 pairs its callers use)."""
    width = parameters["burst_width"][component]
    dm = parameters["dm"][0]
    dm_index = parameters["dm_index"][0]
    ref_freq = parameters["ref_freq"][component]
    sc_time = parameters["scattering_timescale"][0]
    sc_time_freq = sc_time * (freq / ref_freq) ** parameters["scattering_index"][0]
    pair = (name1, name2)
    if pair == ("arrival_time", "burst_width"):
        return 1 / width ** 2 / SQ2
    if pair == ("burst_width", "burst_width"):
        return SQ2 * timediff / width ** 3
    if pair in (("burst_width", "dm"), ("burst_width", "dm_index")):
        return -_deriv_time_dm(name2, freq, ref_freq, dm, dm_index) / width ** 2 / SQ2
    if pair == ("burst_width", "scattering_index"):
        return np.log(freq / ref_freq) / SQ2 / sc_time_freq
    if pair == ("burst_width", "scattering_timescale"):
        return 1 / sc_time / sc_time_freq / SQ2
    if pair == ("scattering_timescale", "scattering_timescale"):
        return -SQ2 * width / sc_time_freq / sc_time ** 2
    if pair in (("arrival_time", "arrival_time"), ("arrival_time", "scattering_timescale"),
                ("arrival_time", "dm"), ("dm", "dm"), ("dm", "scattering_timescale")):
        return 0.0
    raise NameError("sys")  # the reference calls sys.exit without importing sys (derivative.py:450)


def _deriv2_amplitude_pbf_alpha_alpha(freq, parameters, component, normalize):
    """ref: derivative.py:154-209, the only pair requested by a caller (:2630)."""
    width = parameters["burst_width"][component]
    ref_freq = parameters["ref_freq"][component]
    sc_index = parameters["scattering_index"][0]
    sc_time = parameters["scattering_timescale"][0]
    if normalize:
        sc_time_freq = sc_time * (freq / ref_freq) ** sc_index
        return np.sqrt(np.pi / 2) * np.log(freq / ref_freq) ** 2 * width / sc_time_freq
    return np.log(freq / ref_freq) ** 2 * (freq / ref_freq) ** -sc_index


class OracleSecondDerivatives:
    """deriv2_model_wrt_<a>_<b>(parameters, model, component) as methods `second(a, b, ...)`."""

    def __init__(self, model):
        self.model = model
        self.first = OracleDerivatives(model)
        self.normalize = model.normalize_pbf

    # -- shared pieces ---------------------------------------------------------------------
    def _ctx(self, parameters, comp):
        m = self.model
        f = m.freqs[:, None]
        ref_freq = parameters["ref_freq"][comp]
        ctx = dict(
            f=f, w=parameters["burst_width"][comp], ref_freq=ref_freq,
            M=m.spectrum_per_component[:, :, comp], T=m.timediff_per_component[:, :, comp],
            A=m.amplitude_per_component[:, :, comp], L=np.log(f / ref_freq),
            tau0=parameters["scattering_timescale"][0], dm=parameters["dm"][0], beta=parameters["dm_index"][0],
        )
        ctx["tau"] = ctx["tau0"] * (f / ref_freq) ** parameters["scattering_index"][0]
        ctx["apbf"] = _amplitude_pbf(f, parameters, comp, self.normalize)
        with np.errstate(all="ignore"):
            ctx["G"] = ctx["A"] * ctx["apbf"] * np.exp(_argument_exp(ctx["w"], ctx["T"], ctx["tau"])) * 2 / np.sqrt(np.pi)
        return ctx

    def _d1(self, name, parameters, comp):
        return self.first.first(name, parameters, comp, add_all=False)

    def _dae(self, name, c, parameters, comp):
        return _deriv_argument_erf(name, c["f"], c["T"], parameters, comp)

    def _dax(self, name, c, parameters, comp):
        return _deriv_argument_exp(name, c["f"], c["T"], parameters, comp)

    def _d2ae(self, n1, n2, c, parameters, comp):
        return _deriv2_argument_erf(n1, n2, c["f"], c["T"], parameters, comp)

    # -- the 45 functions ------------------------------------------------------------------
    def second(self, name1, name2, parameters, component=0):
        """Dispatch on the reference's function name deriv2_model_wrt_<name1>_<name2>."""
        with np.errstate(all="ignore"):
            return self._second(name1, name2, parameters, component)

    def _second(self, n1, n2, P, comp):
        m = self.model
        M = m.spectrum_per_component[:, :, comp]
        L = np.log(m.freqs / P["ref_freq"][comp])[:, None]
        # rows that are scalar multiples of the model or of a first derivative (:938-1573)
        scalar_rows = {"amplitude": np.log(10), "spectral_running": L ** 2, "spectral_index": L}
        if n1 in scalar_rows:
            factor = scalar_rows[n1]
            if n2 == "amplitude":
                return np.log(10) ** 2 * M
            if n1 == "amplitude" and n2 == "spectral_running":
                return np.log(10) * L ** 2 * M
            if n1 == "amplitude" and n2 == "spectral_index":
                return np.log(10) * L * M
            if n1 == "spectral_running" and n2 == "spectral_running":
                return L ** 4 * M
            if n1 == "spectral_running" and n2 == "spectral_index":
                return L ** 3 * M
            if n1 == "spectral_index" and n2 == "spectral_index":
                return L ** 2 * M
            return factor * self._d1(n2, P, comp)
        if n1 == "burst_width":
            return self._burst_width_row(n2, P, comp)
        if n1 == "arrival_time":
            return self._arrival_time_row(n2, P, comp)
        if (n1, n2) == ("dm", "dm_index"):
            return self._dm_dm_index(P, comp)
        # everything below returns the SUM over components regardless of `comp`
        total = 0.0
        for cur in range(m.spectrum_per_component.shape[2]):
            total = total + self._global_pair(n1, n2, P, cur, comp)
        return total

    def _burst_width_row(self, n2, P, comp):
        c = self._ctx(P, comp)
        w, M, T, tau, G = c["w"], c["M"], c["T"], c["tau"], c["G"]
        gauss = (tau == 0.0)
        d1 = self._d1(n2, P, comp)
        dae_w = self._dae("burst_width", c, P, comp)
        if n2 == "burst_width":  # :1575-1637
            g = -3 * T ** 2 * M / w ** 4 + T ** 4 * M / w ** 6
            s = M / tau ** 2 + w * d1 / tau ** 2 + G * self._d2ae("burst_width", "burst_width", c, P, comp) + \
                G * dae_w * self._dax("burst_width", c, P, comp)
            return np.where(gauss, g, s)
        if n2 == "arrival_time":  # :1639-1700
            g = -2 * T * M / w ** 3 + T ** 3 * M / w ** 5
            s = w * d1 / tau ** 2 + G * self._d2ae("arrival_time", "burst_width", c, P, comp) + \
                G * dae_w * self._dax("arrival_time", c, P, comp)
            return np.where(gauss, g, s)
        if n2 == "scattering_timescale":  # :1702-1759 (no Gaussian branch)
            return -2 * w * M / c["tau0"] / tau ** 2 + w * d1 / tau ** 2 + \
                G * self._d2ae("burst_width", "scattering_timescale", c, P, comp) + \
                G * dae_w * self._dax("scattering_timescale", c, P, comp)
        if n2 == "scattering_index":  # :1761-1820
            return -2 * c["L"] * w * M / tau ** 2 + w * d1 / tau ** 2 - G * dae_w * c["L"] + \
                G * self._d2ae("burst_width", "scattering_index", c, P, comp) + \
                G * dae_w * self._dax("scattering_index", c, P, comp)
        if n2 in ("dm", "dm_index"):  # :1822-1955 (note the sign flip between the two Gaussian branches)
            dtd = _deriv_time_dm(n2, c["f"], c["ref_freq"], c["dm"], c["beta"])
            sign = 2 if n2 == "dm" else -2
            g = sign * T * dtd * M / w ** 3 + T ** 2 * d1 / w ** 3
            s = w * d1 / tau ** 2 + G * self._d2ae("burst_width", n2, c, P, comp) + G * dae_w * self._dax(n2, c, P, comp)
            return np.where(gauss, g, s)
        raise KeyError(n2)

    def _arrival_time_row(self, n2, P, comp):
        c = self._ctx(P, comp)
        w, M, T, tau, G = c["w"], c["M"], c["T"], c["tau"], c["G"]
        gauss = (tau == 0.0)
        dae_t = self._dae("arrival_time", c, P, comp)
        if n2 == "arrival_time":  # :1957-2014
            d1 = self._d1("arrival_time", P, comp)
            g = T ** 2 * M / w ** 4 - M / w ** 2
            s = d1 / tau + G * 0.0 + G * dae_t * self._dax("arrival_time", c, P, comp)
            return np.where(gauss, g, s)
        if n2 == "dm":  # :2016-2080
            d1 = self._d1("dm", P, comp)
            dtd = _deriv_time_dm("dm", c["f"], c["ref_freq"], c["dm"], c["beta"])
            g = T * d1 / w ** 2 + dtd * M / w ** 2
            s = d1 / tau + G * dae_t * self._dax("dm", c, P, comp)
            return np.where(gauss, g, s)
        if n2 == "dm_index":  # :2082-2115: Gaussian form always, SUMMED first derivative
            summed = self.first.dm_index(P, comp)  # add_all=True
            f, beta, rf = c["f"], c["beta"], c["ref_freq"]
            out = T * summed / w ** 2
            out = out + (DISPERSION_CONSTANT * c["dm"] * (np.log(f) * f ** beta - np.log(rf) * rf ** beta) * M / w ** 2)
            return out
        if n2 == "scattering_timescale":  # :2117-2174
            d1 = self._d1("scattering_timescale", P, comp)
            return -M / c["tau0"] / tau + d1 / tau + G * 0.0 + G * dae_t * self._dax("scattering_timescale", c, P, comp)
        if n2 == "scattering_index":  # :2176-2234 (own expressions for erf/exp arguments)
            d1 = self._d1("scattering_index", P, comp)
            L = c["L"]
            apbf_d = _deriv_amplitude_pbf("scattering_index", c["f"], P, comp, self.normalize)
            erf_arg = (T / w - w / tau) / SQ2
            e1 = -1.0 / (w * SQ2)
            e2 = L * w / (tau * SQ2)
            exp_arg = w ** 2 / 2 / tau ** 2 - T / tau - erf_arg ** 2
            exp_arg_d = -w ** 2 * L / tau ** 2 + T * L / tau - 2 * erf_arg * e2
            t1 = -M * L / tau
            t2 = d1 / tau
            t3 = 2.0 * c["A"] * np.exp(exp_arg) * apbf_d * e1 / np.sqrt(np.pi)
            t4 = 2.0 * c["A"] * np.exp(exp_arg) * c["apbf"] * e1 * exp_arg_d / np.sqrt(np.pi)
            return t1 + t2 + t3 + t4
        raise KeyError(n2)

    def _dm_dm_index(self, P, comp):
        """:2236-2273, including `ref_freq ** 2`."""
        c = self._ctx(P, comp)
        w, M, T, f, beta, rf, dm = c["w"], c["M"], c["T"], c["f"], c["beta"], c["ref_freq"], c["dm"]
        k = DISPERSION_CONSTANT
        summed = self.first.dm_index(P, comp)
        q = np.log(f) * f ** beta - np.log(rf) * rf ** beta
        out = k * (f ** beta - rf ** beta) * T * summed / w ** 2
        out = out + k ** 2 * dm * M / w ** 2 * (f ** beta - rf ** 2) * q
        out = out - k * q * T * M / w ** 2
        return out

    def _global_pair(self, n1, n2, P, cur, outer):
        """Per-component term of the global x global functions, :2275-2836."""
        c = self._ctx(P, cur)
        w, M, T, tau, G, L, tau0 = c["w"], c["M"], c["T"], c["tau"], c["G"], c["L"], c["tau0"]
        f, beta, rf, dm, k = c["f"], c["beta"], c["ref_freq"], c["dm"], DISPERSION_CONSTANT
        pair = (n1, n2)
        if pair == ("dm", "dm"):  # :2275-2341
            d1 = self._d1("dm", P, cur)
            dtd = _deriv_time_dm("dm", f, rf, dm, beta)
            g = (dtd * T) ** 2 * M / w ** 4 - (dtd / w) ** 2 * M
            s = -dtd * d1 / tau + G * 0.0 + G * self._dae("dm", c, P, cur) * self._dax("dm", c, P, cur)
            return np.where(tau == 0.0, g, s)
        if pair == ("scattering_timescale", "scattering_timescale"):  # :2343-2404
            d1 = self._d1("scattering_timescale", P, cur)
            t0 = (T / tau - (w / tau) ** 2) / tau0
            t0d = -(T / tau - (w / tau) ** 2) / tau0 ** 2 + (-T / tau + 2 * (w / tau) ** 2) / tau0 ** 2
            return t0d * M + t0 * d1 + \
                G * self._d2ae("scattering_timescale", "scattering_timescale", c, P, cur) + \
                G * self._dae("scattering_timescale", c, P, cur) * self._dax("scattering_timescale", c, P, cur)
        if pair == ("scattering_timescale", "scattering_index"):  # :2406-2472
            d1 = self._d1("scattering_index", P, cur)
            erf_arg = (T / w - w / tau) / SQ2
            e1 = w / (SQ2 * tau0 * tau)
            e2 = w * L / tau / SQ2
            e12 = -L * e1
            apbf_d2 = _deriv_amplitude_pbf("scattering_index", f, P, cur, self.normalize)
            exp_arg = (w / tau) ** 2 / 2 - T / tau - erf_arg ** 2
            exp_arg_d2 = (-(w / tau) ** 2 * L + T * L / tau - 2 * erf_arg * e2)
            t0 = (-w ** 2 / tau ** 2 / tau0 + T / tau0 / tau)
            t0d2 = (2 * (w / tau) ** 2 / tau0 * L - T * L / tau / tau0)
            E = 2.0 * c["A"] * np.exp(exp_arg) / np.sqrt(np.pi)
            return t0d2 * M + t0 * d1 + E * apbf_d2 * e1 + E * c["apbf"] * e1 * exp_arg_d2 + E * c["apbf"] * e12
        if pair == ("scattering_timescale", "dm"):  # :2474-2535
            d1 = self._d1("scattering_timescale", P, cur)
            dtd = _deriv_time_dm("dm", f, rf, dm, beta)
            return -dtd * d1 / tau + dtd * M / tau0 / tau + G * 0.0 + \
                G * self._dae("dm", c, P, cur) * self._dax("scattering_timescale", c, P, cur)
        q = np.log(f) * f ** beta - np.log(rf) * rf ** beta
        erf_arg = (T / w - w / tau) / SQ2
        exp_arg = (w / tau) ** 2 / 2 - T / tau - erf_arg ** 2
        if pair == ("scattering_timescale", "dm_index"):  # :2537-2594
            d1 = self._d1("dm_index", P, cur)
            t0 = (-1 / tau0 - w ** 2 / tau ** 2 / tau0 + T / tau0 / tau)
            t0d = -k * dm * q / tau / tau0
            ed = -k * dm * q / w / SQ2
            xd = k * dm * q / tau - 2 * erf_arg * ed
            t3 = np.sqrt(2 / np.pi) * c["A"] * c["apbf"] * w * np.exp(exp_arg) * xd / tau / tau0
            return t0d * M + t0 * d1 + t3
        if pair == ("scattering_index", "scattering_index"):  # :2596-2664
            d1 = self._d1("scattering_index", P, cur)
            ad = _deriv_amplitude_pbf("scattering_index", f, P, cur, self.normalize)
            ad2 = _deriv2_amplitude_pbf_alpha_alpha(f, P, cur, self.normalize)
            ed = w * L / tau / SQ2
            ed2 = -w * L ** 2 / tau / SQ2
            xd = -L * (w / tau) ** 2 + T * L / tau - 2 * erf_arg * ed
            t0 = ad / c["apbf"] - (w ** 2 / tau ** 2 - T / tau) * L
            t0d = (ad2 / c["apbf"] - (ad / c["apbf"]) ** 2 - L * (-2 * L * (w / tau) ** 2 + T * L / tau))
            E = 2.0 * c["A"] * np.exp(exp_arg) / np.sqrt(np.pi)
            return t0d * M + t0 * d1 + E * ad * ed + E * c["apbf"] * ed * xd + E * c["apbf"] * ed2
        if pair in (("scattering_index", "dm"), ("scattering_index", "dm_index")):  # :2666-2777
            d1 = self._d1(n2, P, cur)
            fd = (f ** beta - rf ** beta) if n2 == "dm" else dm * q
            t0 = -(1 + w ** 2 / tau ** 2 - T / tau) * L
            t0d = -L * k * fd / tau
            ed = -k * fd / w / SQ2
            xd = k * fd / tau - 2 * erf_arg * ed
            t3 = np.sqrt(2 / np.pi) * c["A"] * c["apbf"] * w * np.exp(exp_arg) * L * xd / tau
            return t0d * M + t0 * d1 + t3
        if pair == ("dm_index", "dm_index"):  # :2779-2836, amp_pbf from the OUTER component
            d1 = self._d1("dm_index", P, cur)
            apbf_outer = _amplitude_pbf(f, P, outer, self.normalize)
            ed = -k * dm * q / w / SQ2
            product = k * dm * q
            product_d = k * dm * (np.log(f) ** 2 * f ** beta - np.log(rf) ** 2 * rf ** beta)
            xd = product / tau - 2 * erf_arg * ed
            t1 = product_d * M / tau
            t2 = product * d1 / tau
            t3 = -np.sqrt(2 / np.pi) * product_d * c["A"] * apbf_outer * np.exp(exp_arg) / w / tau
            t4 = -np.sqrt(2 / np.pi) * product * c["A"] * apbf_outer * np.exp(exp_arg) / w / tau * xd
            return t1 + t2 + t3 + t4
        raise KeyError(pair)


# names for which the reference defines deriv2_model_wrt_<a>_<b> (first name -> second names)
REFERENCE_PAIRS = {
    "amplitude": ["amplitude", "spectral_running", "spectral_index", "burst_width", "arrival_time",
                  "dm", "dm_index", "scattering_timescale", "scattering_index"],
    "spectral_running": ["spectral_running", "spectral_index", "burst_width", "arrival_time", "dm",
                         "dm_index", "scattering_timescale", "scattering_index"],
    "spectral_index": ["spectral_index", "burst_width", "arrival_time", "dm", "dm_index",
                       "scattering_timescale", "scattering_index"],
    "burst_width": ["burst_width", "arrival_time", "scattering_timescale", "scattering_index", "dm", "dm_index"],
    "arrival_time": ["arrival_time", "dm", "dm_index", "scattering_timescale", "scattering_index"],
    "dm": ["dm_index", "dm"],
    "scattering_timescale": ["scattering_timescale", "scattering_index", "dm", "dm_index"],
    "scattering_index": ["scattering_index", "dm", "dm_index"],
    "dm_index": ["dm_index"],
}


def canonical_pair(name1, name2):
    """The (first, second) order under which the reference defines the function (fitter.py:146-150)."""
    if name2 in REFERENCE_PAIRS.get(name1, []):
        return name1, name2
    return name2, name1


def compute_hessian(fitter, data):
    """ref: analysis/fitter.py:81-178. Returns (hessian, labels)."""
    model = fitter.model
    parameters = model.get_parameters_dict()
    residual = data - model.compute_model(data=data)
    second = OracleSecondDerivatives(model)
    labels_out, idxs, names = [], [], []
    for name in fitter.fit_parameters:
        if name in FITTER_GLOBAL_PARAMETERS:
            labels_out.append(name)
            idxs.append(0)
            names.append(name)
        else:
            labels_out += [f"{name}{i + 1}" for i in range(model.num_components)]
            idxs += list(range(model.num_components))
            names += [name] * model.num_components
    n = len(names)
    hessian = np.zeros((n, n))
    w2 = fitter.weights[:, None] ** 2
    firsts = [second.first.first(names[i], parameters, idxs[i]) for i in range(n)]
    for i in range(n):
        for j in range(i, n):
            mixed = 0
            both_local = names[i] not in FITTER_GLOBAL_PARAMETERS and names[j] not in FITTER_GLOBAL_PARAMETERS
            if not (both_local and idxs[i] != idxs[j]):
                comp = idxs[i] if names[j] in FITTER_GLOBAL_PARAMETERS else idxs[j]
                a, b = canonical_pair(names[i], names[j])
                mixed = second.second(a, b, parameters, comp)
            with np.errstate(all="ignore"):
                hessian[i, j] = np.sum(2 * (firsts[i] * firsts[j] - residual * mixed) * w2)
            hessian[j, i] = hessian[i, j]
    return hessian, labels_out
