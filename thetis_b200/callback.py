"""
Device-resident versions of the conservation diagnostics of `thetis.callback`
(thetis/callback.py:301-484): same class names, `variable_names`, `__call__`
return values and `message_str`, but the integrals / extrema are reduced on the
GPU by one fused kernel each (tb_swe_integrals, tb_tracer_integrals) and only
4 doubles cross the bus -- the fields themselves stay on the device (SURVEY.md
8f rank 3).  HDF5 output (`export_to_hdf5`) is outside the accelerated path.
"""
from __future__ import annotations

import numpy as np
import torch

__all__ = ["DiagnosticCallback", "ScalarConservationCallback", "VolumeConservation2DCallback",
           "TracerMassConservation2DCallback", "ConservativeTracerMassConservation2DCallback",
           "MinMaxConservationCallback", "TracerOvershootCallBack"]


def _print(msg):
    print(msg, flush=True)


class DiagnosticCallback:
    """callback.py:155-299 (log output only)"""
    name = "diagnostic"
    variable_names = []

    def __init__(self, solver_obj, outputdir=None, array_dim=1, attrs=None, export_to_hdf5=False, append_to_log=True,
                 include_time=True, hdf5_dtype="d", start_time=None, end_time=None):
        if export_to_hdf5:
            raise NotImplementedError("HDF5 diagnostics output is outside the accelerated path")
        self.solver_obj = solver_obj
        self.append_to_log = append_to_log
        self.start_time = -np.inf if start_time is None else start_time
        self.end_time = np.inf if end_time is None else end_time
        self.history = []                      # (time, values) of every evaluation

    def message_str(self, *args):
        return "{} diagnostic".format(self.name)

    def push_to_log(self, time, args):
        _print(self.message_str(*args))

    def evaluate(self, index=None):
        time = self.solver_obj.simulation_time
        if time < self.start_time or time > self.end_time:
            return
        values = self.__call__()
        self.history.append((time, values))
        if self.append_to_log:
            self.push_to_log(time, values)

    # ---- device helpers
    def _swe(self):
        return self.solver_obj._swe_stepper()

    def _tracer(self, tracer_name):
        ts = self.solver_obj.timestepper
        st = getattr(ts, "timesteppers", {}).get(tracer_name)
        if st is None:
            raise KeyError(f"no B200 integrator owns tracer field {tracer_name!r}")
        return st

    @staticmethod
    def _reduce(stepper, out, ops):
        """global value over all ranks: sums via all-reduce, extrema via min/max"""
        if stepper.halo is not None:
            stepper.halo.allreduce(out, ops)
        return out.cpu().numpy()


class ScalarConservationCallback(DiagnosticCallback):
    """callback.py:301-332"""
    variable_names = ["integral", "relative_difference"]

    def __init__(self, scalar_callback, solver_obj, **kwargs):
        super().__init__(solver_obj, **kwargs)
        self.scalar_callback = scalar_callback
        self.initial_value = None

    def __call__(self):
        value = self.scalar_callback()
        if self.initial_value is None:
            self.initial_value = value
        rel_diff = (value - self.initial_value) / self.initial_value
        return value, rel_diff

    def message_str(self, *args):
        return "{0:s} rel. error {1:11.4e}".format(self.name, args[1])


class VolumeConservation2DCallback(ScalarConservationCallback):
    """Checks conservation of 2D volume = int (elev + bathymetry) dx (callback.py:352-367, utility comp_volume_2d)"""
    name = "volume2d"

    def __init__(self, solver_obj, **kwargs):
        def vol2d():
            sw = self._swe()
            out = torch.zeros(4, dtype=torch.float64, device=sw.engine.device)
            sw.engine.swe_integrals(sw.device_state(), out)
            return float(self._reduce(sw, out, "ssss")[3])
        super().__init__(vol2d, solver_obj, **kwargs)


class TracerMassConservation2DCallback(ScalarConservationCallback):
    """Depth-averaged tracer mass = int H c dx (callback.py:369-389, comp_tracer_mass_2d)"""
    name = "tracer mass"

    def __init__(self, tracer_name, solver_obj, **kwargs):
        self.name = tracer_name + " mass"

        def mass():
            st = self._tracer(tracer_name)
            out = torch.zeros(4, dtype=torch.float64, device=st.engine.device)
            st.engine.tracer_integrals(st.device_state(), st._swe_state_for_tracer(), out)
            return float(self._reduce(st, out, "ssmM")[1])
        super().__init__(mass, solver_obj, **kwargs)


class ConservativeTracerMassConservation2DCallback(ScalarConservationCallback):
    """Depth-integrated (conservative form) tracer mass = int q dx (callback.py:392-411)"""
    name = "tracer mass"

    def __init__(self, tracer_name, solver_obj, **kwargs):
        self.name = tracer_name + " mass"

        def mass():
            st = self._tracer(tracer_name)
            out = torch.zeros(4, dtype=torch.float64, device=st.engine.device)
            st.engine.tracer_integrals(st.device_state(), st._swe_state_for_tracer(), out)
            return float(self._reduce(st, out, "ssmM")[0])
        super().__init__(mass, solver_obj, **kwargs)


class MinMaxConservationCallback(DiagnosticCallback):
    """callback.py:431-460"""
    variable_names = ["min_value", "max_value", "undershoot", "overshoot"]

    def __init__(self, minmax_callback, solver_obj, **kwargs):
        super().__init__(solver_obj, **kwargs)
        self.minmax_callback = minmax_callback
        self.initial_value = None

    def __call__(self):
        value = self.minmax_callback()
        if self.initial_value is None:
            self.initial_value = value
        overshoot = max(value[1] - self.initial_value[1], 0.0)
        undershoot = min(value[0] - self.initial_value[0], 0.0)
        return value[0], value[1], undershoot, overshoot

    def message_str(self, *args):
        return "{0:s} {1:g} {2:g}".format(self.name, args[2], args[3])


class TracerOvershootCallBack(MinMaxConservationCallback):
    """Checks overshoots of the given tracer field (callback.py:463-484)"""
    name = "tracer overshoot"

    def __init__(self, tracer_name, solver_obj, **kwargs):
        self.name = tracer_name + " overshoot"

        def minmax():
            st = self._tracer(tracer_name)
            out = torch.zeros(4, dtype=torch.float64, device=st.engine.device)
            st.engine.tracer_integrals(st.device_state(), st._swe_state_for_tracer(), out)
            o = self._reduce(st, out, "ssmM")
            return float(o[2]), float(o[3])
        super().__init__(minmax, solver_obj, **kwargs)
