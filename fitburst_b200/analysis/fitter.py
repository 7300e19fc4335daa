"""
LSFitter: drop-in mirror of fitburst.analysis.fitter.LSFitter (reference analysis/fitter.py:16-542).

fit() runs the same algorithm the reference gets from scipy.optimize.least_squares defaults
(trust-region-reflective without bounds), restated on the normal equations in
fitburst_b200/csrc/fb_trf.h and fed by one fused CUDA pass per trial point that returns chi^2,
J^T r and J^T J without ever materialising the Jacobian; the exact Hessian and the
uncertainties come from a second fused pass. Same attributes, methods, print-outs, error
convention ("fit() never raises") and fit_statistics keys as the reference.
"""
import ctypes
import traceback
import weakref

import numpy as np
import torch
from scipy.optimize import OptimizeResult, least_squares

from .. import _lib

from .._say import quiet, say as _say  # thread-local silencing, see fitburst_b200/_say.py


_TERMINATION_MESSAGES = {  # scipy/optimize/_lsq/least_squares.py TERMINATION_MESSAGES
    -1: "Improper input parameters status returned from `leastsq`",
    0: "The maximum number of function evaluations is exceeded.",
    1: "`gtol` termination condition is satisfied.",
    2: "`ftol` termination condition is satisfied.",
    3: "`xtol` termination condition is satisfied.",
    4: "Both `ftol` and `xtol` termination conditions are satisfied.",
}


class LazyOptimizeResult(OptimizeResult):
    """
    OptimizeResult whose large members `fun` (m,) and `jac` (m, n) are produced on first access:
    the fit itself only ever forms J^T J, J^T r and chi^2 (2.3 GB of Jacobian at 16384 x 1024 x 17).
    """

    def __init__(self, *args, lazy=None, **kwargs):
        super().__init__(*args, **kwargs)
        object.__setattr__(self, "_lazy", lazy or {})

    def __getitem__(self, key):
        lazy = object.__getattribute__(self, "_lazy")
        if key in lazy and not dict.__contains__(self, key):
            dict.__setitem__(self, key, lazy[key]())
        return dict.__getitem__(self, key)

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as exc:
            raise AttributeError(name) from exc


class LSFitter:
    """
    A Python object that defines methods and configurations for
    least-squares fitting of radio dynamic spectra.
    """

    # pylint: disable=too-many-instance-attributes

    def __init__(self, data: float, model: object, good_freq: bool, weighted_fit: bool = True,
                 weight_range: list = None):
        """Same arguments as the reference constructor (analysis/fitter.py:22-79)."""
        self.data = data
        self.fit_parameters = []
        self.fit_statistics = {}
        self.model = model

        if self.model.scintillation:  # analysis/fitter.py:60-65
            for current_parameter in self.model.parameters:
                if current_parameter not in ["amplitude", "spectral_index", "spectral_running"]:
                    self.fit_parameters += [current_parameter]
        else:
            self.fit_parameters = self.model.parameters.copy()

        self.good_freq = good_freq
        self.global_parameters = ["dm", "dm_index", "scattering_timescale", "scattering_index"]
        self.success = None
        self.weights = None
        self.weighted_fit = weighted_fit
        self.weight_range = weight_range

        self._data_dev = None
        self._weights_loaded = None
        self._weights_token = object()  # identifies this fitter's weights on the model's plan
        self._chisq_initial = None
        self._set_weights()

    # ------------------------------------------------------------------ native plumbing
    def _lib(self):
        return _lib.load_library()

    def _stream(self):
        return self.model._stream()

    def _data_on_device(self):
        if self._data_dev is None:
            self._data_dev = self.model._to_device(self.data)
        return self._data_dev

    def _sync_weights(self):
        """Pushes self.weights to the plan when they are not the ones it holds: user code replaced
        the array, the plan was rebuilt, or ANOTHER fitter on the same model loaded its own since
        (the weights and the good-channel list live on the model's plan; in the reference every
        fitter owns its weights, analysis/fitter.py:506-542, so two live fitters on one model
        must not see each other's)."""
        plan = self.model._ensure_plan()
        weights = np.ascontiguousarray(self.weights, dtype=np.float64)
        changed = self._weights_loaded is None or \
            not np.array_equal(weights, self._weights_loaded, equal_nan=True)
        if changed and self._weights_loaded is not None:
            self._chisq_initial = None  # stale: recomputed from the current weights when needed
        if changed or self._weights_plan is not plan or \
                self.model.__dict__.get("_weights_owner") is not self._weights_token:
            with torch.cuda.device(self.model._device):
                _lib.check(self._lib().fb_load_weights(plan, _lib.dptr(weights), self._stream()))
            self._weights_loaded = weights.copy()
            self._weights_plan = plan
            self.model.__dict__["_weights_owner"] = self._weights_token
        return plan

    def _sync_free_parameters(self, plan):
        mask = (ctypes.c_int32 * 9)()
        for idx, name in enumerate(self.model.parameters):
            mask[idx] = 1 if name in self.fit_parameters else 0
        num_free = ctypes.c_int32(0)
        _lib.check(self._lib().fb_set_free_parameters(plan, mask, ctypes.byref(num_free)))
        return num_free.value

    def _prepare(self):
        plan = self._sync_weights()
        num_free = self._sync_free_parameters(plan)
        return plan, num_free

    # ------------------------------------------------------------------ reference surface
    def compute_hessian(self, data: float, parameter_list: list) -> float:
        """
        Computes the exact chi-squared Hessian at the model's current state
        (analysis/fitter.py:81-178). Returns (hessian, labels).
        """
        _say("INFO: computing hessian matrix with best-fit parameters.")
        cached = getattr(self, "_hessian_native", None)
        if cached is not None and data is self.data and cached[1] == tuple(self.fit_parameters) and \
                cached[0] == self.model._packed_parameters().tobytes() and \
                np.array_equal(np.ascontiguousarray(self.weights, dtype=np.float64), self._weights_loaded, equal_nan=True):
            return cached[2].copy(), self._parameter_labels()  # computed by fb_fit at this very state
        plan, num_free = self._prepare()
        self.model._push_parameters()
        data_dev = self._data_on_device() if data is self.data else self.model._to_device(data)
        hessian = np.zeros((num_free, num_free), dtype=np.float64)
        with torch.cuda.device(self.model._device):
            _lib.check(self._lib().fb_hessian(plan, data_dev.data_ptr(), _lib.dptr(hessian), self._stream()))
        return hessian, self._parameter_labels()

    def _parameter_labels(self):
        """Labels of analysis/fitter.py:116-127."""
        labels = []
        for current_par in self.fit_parameters:
            if current_par in self.global_parameters:
                labels += [current_par]
            else:
                labels += [f"{current_par}{idx+1}" for idx in range(self.model.num_components)]
        return labels

    def compute_jacobian(self, parameter_list: list) -> float:
        """
        Computes the Jacobian matrix of the residual vector at the model's CURRENT state, like the
        reference, which ignores the values of `parameter_list` (analysis/fitter.py:180-229).
        """
        plan, num_free = self._prepare()
        self.model._push_parameters()
        num_points = len(self.model.times) * len(self.model.freqs)
        jac = torch.empty((num_points, num_free), dtype=torch.float64, device=self.model._device)
        with torch.cuda.device(self.model._device):
            _lib.check(self._lib().fb_jacobian_dense(
                plan, self._data_on_device().data_ptr(), jac.data_ptr(), self._stream()))
        return jac.cpu().numpy()

    def compute_residuals(self, parameter_list: list) -> float:
        """
        Computes the weighted residual vector (analysis/fitter.py:231-266); updates the model.
        """
        parameter_dict = self.load_fit_parameters_list(parameter_list)
        self.model.update_parameters(parameter_dict)
        plan, _ = self._prepare()
        self.model._push_parameters()
        if self.model.verbose:
            self.model._print_parameters()
        resid = torch.empty(self.model.num_freq * self.model.num_time, dtype=torch.float64,
                            device=self.model._device)
        with torch.cuda.device(self.model._device):
            _lib.check(self._lib().fb_residuals(
                plan, self._data_on_device().data_ptr(), resid.data_ptr(), self._stream()))
        self.model._cache_state = (self.model._packed_parameters(),
                                   self._data_on_device() if self.model.scintillation else None)
        self.model._cache_host = {}
        return resid.cpu().numpy()

    def fit(self, exact_jacobian: bool = True) -> None:
        """
        Executes least-squares fitting of the model spectrum to data, and stores
        results to child class (analysis/fitter.py:268-322). Never raises.
        """
        # pylint: disable=broad-except
        parameter_list = self.get_fit_parameters_list()
        try:
            if exact_jacobian:
                results = self._fit_native(parameter_list)
            else:
                # compatibility path: scipy's own '2-point' finite differences around the CUDA
                # residual function, exactly what the reference does for exact_jacobian=False.
                results = least_squares(self.compute_residuals, parameter_list, jac="2-point")
            self.results = results

            if self.results.success:
                _say("INFO: fit successful!")
            else:
                _say("INFO: fit didn't work!")

            self._compute_fit_statistics(self.data, results)

            if self.success:
                _say("INFO: derived uncertainties and fit statistics")

        except Exception as exc:
            _say("ERROR: solver encountered a failure! Debug!")
            _say(traceback.format_exc())

    def _fit_native(self, parameter_list):
        """The trust-region loop of fb_fit() wrapped into a scipy-style OptimizeResult."""
        plan, num_free = self._prepare()
        if num_free != len(parameter_list):
            raise ValueError("free-parameter layout mismatch between model and fitter")
        self.model._push_parameters()
        options = _lib.FitOptions(1e-8, 1e-8, 1e-8, 0, 1)
        result = _lib.FitResult()
        gram = np.zeros((num_free + 1, num_free + 1), dtype=np.float64)
        hessian = np.zeros((num_free, num_free), dtype=np.float64)
        packed = np.zeros((_lib.FB_NUM_PARAMETERS, int(self.model.num_components)), dtype=np.float64)
        self._hessian_native = None
        with torch.cuda.device(self.model._device):
            # one call: device-driven trust-region loop + exact Hessian at the last evaluated point
            _lib.check(self._lib().fb_fit(
                plan, self._data_on_device().data_ptr(), ctypes.byref(options), ctypes.byref(result),
                _lib.dptr(gram), _lib.dptr(hessian), self._stream()))
            # the reference leaves the model at the LAST EVALUATED point (analysis/fitter.py:258);
            # mirror that on the python side from the plan's parameter block.
            _lib.check(self._lib().fb_get_parameters(plan, _lib.dptr(packed), self._stream()))
        x = np.array(result.x[:num_free], dtype=np.float64)
        last_x = []
        for name in self.model.parameters:
            if name in self.fit_parameters:
                row = packed[_lib.PARAMETER_INDEX[name]]
                last_x += [row[0]] if name in self.global_parameters else row.tolist()
        self.model.update_parameters(self.load_fit_parameters_list(last_x))
        self.model._cache_state = (self.model._packed_parameters(),
                                   self._data_on_device() if self.model.scintillation else None)
        self.model._cache_host = {}
        self._gram_final = gram
        # compute_hessian() right after the fit (what _compute_fit_statistics does) is this matrix
        self._hessian_native = (self.model._packed_parameters().tobytes(), tuple(self.fit_parameters), hessian)

        # the lazy members must not keep the fitter alive through a reference cycle
        # (fitter -> results -> closure -> fitter): dead fitters would hold their 134 MB device
        # spectrum until a cyclic-GC pass, which also stalls the caller for tens of milliseconds.
        owner = weakref.ref(self)

        def lazy_fun():
            return owner()._evaluate_at(x, want="fun")

        def lazy_jac():
            return owner()._evaluate_at(x, want="jac")

        results = LazyOptimizeResult(
            x=x, cost=float(result.cost), grad=np.array(result.grad[:num_free]),
            optimality=float(result.optimality), active_mask=np.zeros_like(x),
            nfev=int(result.nfev), njev=int(result.njev), status=int(result.status),
            lazy={"fun": lazy_fun, "jac": lazy_jac},
        )
        results["message"] = _TERMINATION_MESSAGES[results.status]
        results["success"] = results.status > 0
        return results

    def _evaluate_at(self, x, want):
        """Dense residuals / Jacobian at x without disturbing the model's current state."""
        saved = self.model.get_parameters_dict()
        saved = {key: (list(value) if value is not None else None) for key, value in saved.items()}
        saved_state, saved_cache = self.model._cache_state, self.model._cache_host
        try:
            out = self.compute_residuals(list(x))
            if want == "jac":
                out = self.compute_jacobian(list(x))
        finally:
            self.model.update_parameters({k: v for k, v in saved.items() if v is not None})
            self.model._cache_state, self.model._cache_host = saved_state, saved_cache
        return out

    def fix_parameter(self, parameter_list: list) -> None:
        """
        Updates 'fit_parameters' attributes to remove parameters that will
        be held fixed during least-squares fitting (analysis/fitter.py:324-351).
        """
        _say("INFO: removing the following parameters:", ", ".join(parameter_list))
        for current_parameter in parameter_list:
            self.fit_parameters.remove(current_parameter)
        _say("INFO: new list of fit parameters:", ", ".join(self.fit_parameters))

    def get_fit_parameters_list(self) -> list:
        """Values of the fit parameters in packing order (analysis/fitter.py:353-384)."""
        parameter_list = []
        for current_parameter in self.model.parameters:
            if current_parameter in self.fit_parameters:
                current_sublist = getattr(self.model, current_parameter)
                if current_parameter in self.global_parameters:
                    parameter_list += [current_sublist[0]]
                else:
                    parameter_list += current_sublist
        return parameter_list

    def load_fit_parameters_list(self, parameter_list: list) -> dict:
        """Packed list -> {name: per-component list} (analysis/fitter.py:386-430)."""
        parameter_dict = {}
        current_idx = 0
        for current_parameter in self.model.parameters:
            if current_parameter in self.fit_parameters:
                if current_parameter in self.global_parameters:
                    parameter_dict[current_parameter] = [parameter_list[current_idx]]
                    current_idx += 1
                else:
                    parameter_dict[current_parameter] = parameter_list[
                        current_idx: (current_idx + self.model.num_components)
                    ]
                    current_idx += self.model.num_components
        return parameter_dict

    def _compute_fit_statistics(self, spectrum_observed: float, fit_result: object) -> None:
        """
        Computes and stores a variety of statistics and best-fit results
        (analysis/fitter.py:432-504); same keys, all plain Python types.
        """
        # pylint: disable=broad-except
        num_freq, num_time = self.model.num_freq, self.model.num_time
        num_freq_good = int(np.sum(self.good_freq))
        num_fit_parameters = len(fit_result.x)

        self.fit_statistics["num_freq"] = num_freq
        self.fit_statistics["num_freq_good"] = num_freq_good
        self.fit_statistics["num_fit_parameters"] = num_fit_parameters
        self.fit_statistics["num_observations"] = num_freq_good * int(num_time) - num_fit_parameters
        self.fit_statistics["num_time"] = num_time

        native = isinstance(fit_result, LazyOptimizeResult)
        if native and self._chisq_initial is not None:
            chisq_initial = float(self._chisq_initial)
            chisq_final = float(2.0 * fit_result.cost)  # == sum(fun**2), analysis/fitter.py:466
        else:
            chisq_initial = float(np.sum((self.data * self.weights[:, None]) ** 2))
            chisq_final = float(np.sum(fit_result.fun ** 2))
        chisq_final_reduced = chisq_final / self.fit_statistics["num_observations"]

        self.fit_statistics["chisq_initial"] = chisq_initial
        self.fit_statistics["chisq_final"] = chisq_final
        self.fit_statistics["chisq_final_reduced"] = chisq_final_reduced
        self.fit_statistics["snr"] = float(np.sqrt(chisq_initial - chisq_final))

        self.fit_statistics["bestfit_parameters"] = self.load_fit_parameters_list(
            fit_result.x.tolist())
        self.fit_statistics["bestfit_uncertainties"] = None
        self.fit_statistics["bestfit_covariance"] = None

        try:
            if native:
                hessian_approx = self._gram_final[:num_fit_parameters, :num_fit_parameters].copy()
            else:
                hessian_approx = fit_result.jac.T.dot(fit_result.jac)
            covariance_approx = np.linalg.inv(hessian_approx)
            hessian, par_labels = self.compute_hessian(self.data, self.fit_parameters)
            covariance = np.linalg.inv(0.5 * hessian)
            uncertainties = [float(x) for x in np.sqrt(np.diag(covariance)).tolist()]

            if np.any(np.isnan(uncertainties)):
                _say("WARNING: one or more NaNs in exact uncertainties, replacing with approximates...")
                uncertainties = [float(x) for x in np.sqrt(np.diag(covariance_approx)).tolist()]

            self.covariance_approx = covariance_approx
            self.covariance = covariance
            self.covariance_labels = par_labels
            self.hessian_approx = hessian_approx
            self.hessian = hessian

            self.fit_statistics["bestfit_uncertainties"] = self.load_fit_parameters_list(
                uncertainties)
            self.fit_statistics["bestfit_covariance"] = None  # like analysis/fitter.py:499

        except Exception as exc:
            _say(f"ERROR: {exc}; designating fit as unsuccessful...")
            _say(traceback.format_exc())
            self.success = False

    def _set_weights(self) -> None:
        """
        Sets an attribute containing weights to be applied during least-squares fitting
        (analysis/fitter.py:506-542); the default range [0, Nt-1] is a slice, so the last sample
        is excluded exactly like in the reference.
        """
        idx_range = self.weight_range
        if self.weight_range is None:
            idx_range = [0, len(self.data[0, :]) - 1]
        plan = self.model._ensure_plan()
        good = np.ascontiguousarray(np.asarray(self.good_freq, dtype=bool).astype(np.uint8))
        weights = np.empty(self.model.num_freq, dtype=np.float64)
        chisq_initial = ctypes.c_double(0.0)
        with torch.cuda.device(self.model._device):
            _lib.check(self._lib().fb_set_weights(
                plan, self._data_on_device().data_ptr(),
                good.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), int(bool(self.weighted_fit)),
                int(idx_range[0]), int(idx_range[1]), _lib.dptr(weights), ctypes.byref(chisq_initial),
                self._stream()))
        self.weights = weights
        self._weights_loaded = weights.copy()
        self._weights_plan = plan
        self.model.__dict__["_weights_owner"] = self._weights_token
        self._chisq_initial = chisq_initial.value
