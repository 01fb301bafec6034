"""
B200 drop-in for `thetis.limiter.VertexBasedP1DGLimiter` (thetis/limiter.py:48-198):
same constructor and `.apply(field)`; the centroid projection, vertex min/max
bounds (with exterior-facet means) and the per-cell clamp run as CUDA kernels
(tb_limiter_apply).  Scalar 2-D P1DG fields only, like the reference's 2-D path.
"""
from __future__ import annotations

import numpy as np
import torch

from .adaptor import get_adaptor
from .engine import Engine

__all__ = ["VertexBasedP1DGLimiter"]


class VertexBasedP1DGLimiter:
    def __init__(self, p1dg_space, time_dependent_mesh=True):
        el = p1dg_space.ufl_element()
        assert el.family() in ("Discontinuous Lagrange", "DQ") and el.degree() == 1, \
            "function space must be one of ['Discontinuous Lagrange', 'DQ'] of degree 1"   # limiter.py:65
        if getattr(p1dg_space, "value_size", 1) > 1:
            raise NotImplementedError("vector fields are limited component-wise only on extruded meshes (out of scope)")
        self.P1DG = p1dg_space
        self.adaptor = get_adaptor(p1dg_space.mesh())
        self.engine = self.adaptor.get_engine()
        self.halo = self.adaptor.halo
        # the bounds of a vertex use the centroids of ALL cells around it (limiter.py:109-145): on a distributed mesh
        # the one-deep FACET halo is not enough, the partition must carry the vertex-neighbour ghosts
        if self.halo is not None:
            kind = self.adaptor.mesh.meta.get("halo") or getattr(self.halo.part, "halo", None)
            if kind != "vertex":
                raise NotImplementedError(
                    "VertexBasedP1DGLimiter on a distributed mesh needs distribute_mesh(..., halo='vertex'); this "
                    f"partition was built with halo={kind!r}")
        self.node_map = None

    def apply(self, field):
        """Applies the limiter on the given field (in place)."""
        assert field.function_space() is self.P1DG or field.function_space().ufl_element().degree() == 1
        eng = self.engine
        ent = getattr(eng, "tracer_steppers", {}).get(id(field))
        st = ent[1] if ent is not None and ent[0]() is field else None
        if st is not None:
            # the field lives on the device: limit there, no host round trip
            if not st._host_stale and st._host_changed():
                st.upload()
            # out of place into the integrator's scratch buffer (neighbouring patches read each other's ORIGINAL
            # values), then the two buffers swap roles: no copy back
            src, dst = st.buf[0], st.buf[1]
            if self.halo is not None:
                self.halo.limiter_apply_to(src, dst)       # + limited ghost values for the next step
            else:
                eng.limiter_apply_to(src, dst)
            st.buf[0], st.buf[1] = dst, src
            st.mark_device_modified()
            if st.sync_policy == "every_step":
                st.sync_to_host()
            return
        if self.halo is not None:
            raise NotImplementedError("distributed limiter needs the field to be owned by a B200 tracer integrator")
        if self.node_map is None:
            self.node_map = torch.as_tensor(self.adaptor.dg_node_map(self.P1DG).reshape(-1)).to(eng.device)
        q = torch.as_tensor(np.ascontiguousarray(np.asarray(self.adaptor.dat_ro(field), dtype=np.float64))).to(eng.device)
        c = eng.new_tracer()
        eng.tracer_from_field(q, self.node_map, c)
        eng.limiter_apply(c)
        eng.tracer_to_field(c, self.node_map, q)
        self.adaptor.dat_rw(field)[...] = q.cpu().numpy()
