"""User-level fused device ops built on the reference's STATEFUL op protocol
(`create_stateful_op_func` / `OpClass`, reference ops/wrapping.py:47-76,181-217 -- exported by the
reference but unused in its repo).

    y = md.relu(h)                      # ONE launch instead of greater + where (SURVEY finding 2)
    y = md.linear_relu(X, W, b)         # relu(X @ W + b): ONE tcgen05 GEMM launch, bias and ReLU in its epilogue
    y = md.linear(X, W, b)              # X @ W + b in one launch (no ReLU: the output layer)

An op instance lives for one forward call and keeps what its gradient functions share -- the
output `y` (whose sign IS the ReLU mask) and the masked upstream gradient G = grad * (y > 0),
computed once and used by all three of dX = G @ W^T, dW = X^T @ G, db = sum(G, axis=0).

The classes are written against the public `md` API only, so the same code runs under the
UNMODIFIED reference engine with the B200 backend plugged in (`make_ops(reference_md)`), where the
gradient functions are called one by one by the reference's OpNode.update_grads; under this repo's
engine a `fused_backward` handler additionally (a) accumulates into privately owned gradient buffers
through the GEMM epilogue and (b) applies the mask of the layer BELOW inside the dX GEMM's epilogue
when X is itself the output of a linear_relu consumed only here, so that layer's backward starts
from an already-masked gradient and skips its own masking pass.

Results: the forward is bit-identical to the reference composition relu(X @ W + b) with
relu = where(h > 0, h, 0); gradients differ from it only by the GEMM/reduction summation order.
Higher-order sweeps (allow_higher_order=True) run the gradient functions with recording on; they
are then expressed in differentiable `md` ops, so second-order graphs through these ops work.
"""
from __future__ import annotations

from minidiff_b200.backend import functions as F
from minidiff_b200.backend.device_array import F32, DeviceArray


def _raw(t):
    return t._data if hasattr(t, "_data") else t


def _is_f32_matrix(a) -> bool:
    return isinstance(a, DeviceArray) and a.dtype == F32 and a.ndim == 2


def make_ops(md):
    """Build (relu, linear_relu, linear) for engine module `md` (this repo's or the reference's)."""
    wrapping = md.ops.wrapping if hasattr(md, "ops") else md
    OpClass = wrapping.OpClass
    create = wrapping.create_stateful_op_func

    def recording():
        return md.grad_allowed_()

    class Relu(OpClass):
        """y = where(x > 0, x, 0); dx = grad * (x > 0)."""

        def __init__(self):
            self.y = None

        def create_forward(self):
            def forward(x):
                xd = _raw(x)
                if isinstance(xd, DeviceArray) and xd.dtype.kind == "f":
                    y = md.Tensor(F._ew("RELU", xd.dtype, xd))
                else:
                    y = md.where(x > 0, x, 0).detach()
                self.y = y.detach()      # an alias without op_node: no reference cycle through the node
                return y

            return forward

        def create_grads(self):
            def grad_x(x, grad):
                if recording():
                    return grad * (self.y > 0)
                return md.Tensor(F._ew("RELU_MASK_BWD", F.result_dtype(_raw(grad), _raw(self.y)), _raw(grad), _raw(self.y)))

            return [grad_x]

    class LinearRelu(OpClass):
        """y = relu(X @ W + b) (apply_relu=False: y = X @ W + b)."""

        apply_relu = True

        def __init__(self):
            self.y = None
            self._g = None          # masked upstream gradient, shared by the three gradient functions
            self._g_src = None
            self._premasked = None  # a gradient Tensor the layer above already masked with (y > 0)

        # ---------------------------------------------------------------- forward
        def create_forward(self):
            def forward(X, W, b):
                xd, wd, bd = _raw(X), _raw(W), _raw(b)
                out = None
                if _is_f32_matrix(xd) and _is_f32_matrix(wd) and isinstance(bd, DeviceArray) and bd.dtype == F32:
                    out = F._gemm_fused(xd, wd, bias=bd if bd.is_c_contiguous() else F.copy_(bd),
                                        relu=self.apply_relu)
                if out is not None:
                    y = md.Tensor(out)
                else:                      # shapes the pair kernel does not take: the reference chain
                    with md.no_grad():
                        h = md.matmul(X, W) + b
                        y = md.where(h > 0, h, 0) if self.apply_relu else h
                    y = y.detach()
                self.y = y.detach()        # alias without op_node (the node holds this instance: no cycle)
                return y

            return forward

        # ---------------------------------------------------------------- shared state
        def masked(self, grad):
            """G = grad * (y > 0), one launch, computed once per upstream gradient."""
            if not self.apply_relu or grad is self._premasked:
                return grad
            if self._g is None or self._g_src is not grad:
                if recording():
                    g = grad * (self.y > 0)
                else:
                    g = md.Tensor(F._ew("RELU_MASK_BWD", F32, _raw(grad), _raw(self.y)))
                self._g, self._g_src = g, grad
            return self._g

        def create_grads(self):
            def grad_X(X, W, b, grad):
                return md.matmul(self.masked(grad), W.T)

            def grad_W(X, W, b, grad):
                return md.matmul(X.T, self.masked(grad))

            def grad_b(X, W, b, grad):
                return md.sum(self.masked(grad), axis=(0,))   # tuple: sum_grad rejects an int axis (reference quirk)

            return [grad_X, grad_W, grad_b]

        # ---------------------------------------------------------------- this repo's engine only
        def fused_backward(self, node, index, op_input, grad):
            from minidiff_b200.ops import fused as fb
            from minidiff_b200.topology import OpNode

            X, W, b = node.op_inputs
            g = self.masked(grad)
            gd = _raw(g)
            if not _is_f32_matrix(gd) or op_input._data.dtype != F32:
                return False
            if index == 2:
                return fb.contribute(op_input, "COPY", gd)
            if not (_is_f32_matrix(_raw(X)) and _is_f32_matrix(_raw(W))):
                return False
            a, bm = (gd, _raw(W).T) if index == 0 else (_raw(X).T, gd)
            if (a.shape[0], bm.shape[1]) != op_input._data.shape:
                return False
            dst = OpNode.private_grad_buffer(op_input)
            if dst is not None and not (dst.dtype == F32 and dst.is_c_contiguous()):
                dst = None
            if index == 0:
                # X produced by a linear_relu that only this node consumes: fold ITS mask (X > 0)
                # into the epilogue of this GEMM; its backward then starts from a masked gradient
                below = getattr(getattr(X.op_node, "fused_backward", None), "__self__", None)
                if (isinstance(below, LinearRelu) and below.apply_relu and X.graph_refs == 1
                        and dst is None and op_input.grad is None and below.y._data is X._data):
                    out = F._gemm_fused(a, bm, mask_src=X._data)
                    if out is not None:
                        t = md.Tensor(out)
                        below._premasked = t
                        OpNode.accumulate(op_input, t, private=True)
                        return True
            if dst is not None:
                F._gemm(a, bm, out=dst, accumulate=True)
                return True
            OpNode.accumulate(op_input, md.Tensor(F._gemm(a, bm)), private=True)
            return True

    class Linear(LinearRelu):
        apply_relu = False

    relu = create(Relu, tensor_only=True, op_name="relu")
    linear_relu = create(LinearRelu, tensor_only=True, op_name="linear_relu")
    linear = create(Linear, tensor_only=True, op_name="linear")
    return relu, linear_relu, linear
