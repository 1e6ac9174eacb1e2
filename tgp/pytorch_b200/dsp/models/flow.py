"""Flow modules G of the transformed GP — the hot subset of the reference's flow zoo
(reference code/dsp/models/flow.py: CompositeFlow :146-191, IdentityFlow :296-307, AffineFlow :310-361,
TanhFlow :619-815, Sinh_ArcsinhFlow :817-996, StepFlow :1039-1128, switch_off :1130-1149, instance_flow :39-85).

Each module keeps the reference's constructor signature, parameter names (`a`, `b`, `c`, `d`, `NNets_a`, `NNets_b`,
`flow_arr`, `switch_off`) and `forward(f0, X=None)`, so initialisers and optimiser filters written against the
reference keep working.  `forward` is a few torch ops for callers outside the hot path (flow initialisers, sampling).
On the ELBO / test-NLL path the modules are never called: `describe()` hands their structure and parameters to the
fused epilogue kernel (tgp_ell_forward / tgp_test_rows), which evaluates G, G' and dG/dtheta at every quadrature node.
"""
import torch
import torch.nn as nn
from torch.nn.functional import softplus

from .. import config as cg
from ..nn_layers import apply_linear
from ..utils import inv_softplus


def instance_flow(flow_list, is_composite=True):
    built = []
    for name, init in flow_list:
        if name == 'affine':
            fl = AffineFlow(**init)
        elif name == 'sinh_arcsinh':
            fl = Sinh_ArcsinhFlow(**init)
        elif name == 'identity':
            fl = IdentityFlow()
        elif name == 'tanh':
            fl = TanhFlow(**init)
        elif name == 'step_flow':
            fl = StepFlow(**init)
        elif name == 'arcsinh':
            fl = ArcsinhFlow(**init)
        elif name == 'boxcox':
            fl = BoxCoxFlow(**init)
        elif name in ('inverseboxcox', 'inverse_boxcox'):
            fl = InverseBoxCoxFlow(**init)
        else:
            raise ValueError('flow %r is outside the fused hot-path subset (identity, affine, tanh, sinh_arcsinh, step_flow, '
                             'arcsinh, boxcox, inverseboxcox); see DESIGN.md "out of scope"' % (name,))
        built.append(fl)
    return CompositeFlow(built) if is_composite else built


class Flow(nn.Module):
    def forward(self, f0, X=None):
        raise NotImplementedError

    def KLD(self):
        return 0.0

    def forward_initializer(self, X):
        raise NotImplementedError

    def turn_off_initializer_parameters(self):
        raise NotImplementedError

    def describe(self, X=None, n_mc=1):
        """Returns (layer dicts, [global scalar tensors], [per-row tensors of shape (n_mc*MB,)]) for the fused kernels."""
        raise NotImplementedError


class CompositeFlow(Flow):
    def __init__(self, flow_arr):
        super().__init__()
        self.flow_arr = nn.ModuleList(flow_arr)

    def forward(self, f, X=None):
        for fl in self.flow_arr:
            f = fl.forward(f, X)
        return f

    def forward_initializer(self, X):
        return sum(fl.forward_initializer(X) for fl in self.flow_arr)

    def append_flow(self, flows):
        self.flow_arr.extend(flows)

    def KLD(self):
        return sum(fl.KLD() for fl in self.flow_arr)

    @property
    def input_dependent(self):
        return None

    @input_dependent.setter
    def input_dependent(self, value):
        for fl in self.flow_arr:
            fl.input_dependent = value

    def turn_off_initializer_parameters(self):
        for fl in self.flow_arr:
            fl.turn_off_initializer_parameters()

    def describe(self, X=None, n_mc=1):
        layers, glob, rows = [], [], []
        for fl in self.flow_arr:
            l, g, r = fl.describe(X, n_mc)
            layers += l
            glob += g
            rows += r
        return layers, glob, rows


class IdentityFlow(Flow):
    def __init__(self):
        super().__init__()

    def forward(self, f0, X=None):
        return f0

    def inverse(self, f):
        return f

    def forward_initializer(self, X):
        return 0.0

    def turn_off_initializer_parameters(self):
        pass

    def describe(self, X=None, n_mc=1):
        return [], [], []


class AffineFlow(Flow):
    """fk = a*f0 + b (a -> softplus(a) under set_restrictions)."""

    def __init__(self, init_a, init_b, set_restrictions, input_dependent=False, input_dim=-1, input_dependent_config={}):
        super().__init__()
        self.a = nn.Parameter(torch.tensor(init_a, dtype=cg.dtype))
        self.b = nn.Parameter(torch.tensor(init_b, dtype=cg.dtype))
        self.set_restrictions = set_restrictions
        self.input_dependent = input_dependent
        self._input_dependent = input_dependent

    def forward(self, f0, X=None):
        if self._input_dependent and self.input_dependent:
            raise NotImplementedError()
        a = softplus(self.a) if self.set_restrictions else self.a
        return a * f0 + self.b

    def inverse(self, f):
        a = softplus(self.a) if self.set_restrictions else self.a
        return (f - self.b) / a

    def forward_initializer(self, X):
        return 0.0

    def turn_off_initializer_parameters(self):
        pass

    def describe(self, X=None, n_mc=1):
        if self._input_dependent and self.input_dependent:
            raise NotImplementedError()
        return [dict(kind='affine', restrict=bool(self.set_restrictions))], [self.a, self.b], []


class ArcsinhFlow(Flow):
    """fk = a + b*asinh((f0-c)/d) [+ f0]; b, d -> softplus under set_restrictions (reference flow.py:495-557)."""

    def __init__(self, init_a, init_b, init_c, init_d, add_init_f0, set_restrictions):
        super().__init__()
        for n, val in zip('abcd', (init_a, init_b, init_c, init_d)):
            setattr(self, n, nn.Parameter(torch.tensor(val, dtype=cg.dtype)))
        self.set_restrictions = True if add_init_f0 else set_restrictions
        self.add_init_f0 = add_init_f0

    def asinh(self, f):
        return torch.log(f + (f ** 2 + 1) ** 0.5)

    def forward(self, f0, X=None):
        b, d = (softplus(self.b), softplus(self.d)) if self.set_restrictions else (self.b, self.d)
        fk = self.a + b * self.asinh((f0 - self.c) / d)
        return fk + f0 if self.add_init_f0 else fk

    def inverse(self, f):
        b, d = (softplus(self.b), softplus(self.d)) if self.set_restrictions else (self.b, self.d)
        return self.c + d * torch.sinh((f - self.a) / b)

    def forward_initializer(self, X):
        return 0.0

    def turn_off_initializer_parameters(self):
        pass

    def describe(self, X=None, n_mc=1):
        return ([dict(kind='arcsinh', restrict=bool(self.set_restrictions), add_f0=bool(self.add_init_f0))],
                [self.a, self.b, self.c, self.d], [])


class BoxCoxFlow(Flow):
    """fk = (sgn(f0)|f0|^lam - 1)/lam [+ f0]; lam optionally through a user constraint (reference flow.py:377-421)."""

    def __init__(self, init_lam, add_init_f0, constraint=None):
        super().__init__()
        self.lam = nn.Parameter(torch.tensor(init_lam, dtype=cg.dtype))          # shape as given: () or (1,)
        self.add_init_f0 = add_init_f0
        self.constraint = constraint

    def transform_param(self):
        if self.constraint is None:
            lam = self.lam
            if lam == 0:
                lam = lam + 1e-11
        else:
            lam = self.constraint(self.lam)
        assert lam != 0, 'Invalid value for Box Cox parameter. This flow is not defined for values of lam == 0'
        return lam

    def forward(self, f0, X=None):
        lam = self.transform_param()
        sgn = torch.sign(f0)
        fk = (sgn * torch.pow(sgn * f0, lam) - 1) / lam
        return fk + f0 if self.add_init_f0 else fk

    def forward_initializer(self, X):
        return 0.0

    def turn_off_initializer_parameters(self):
        pass

    _kind = 'boxcox'

    def describe(self, X=None, n_mc=1):
        # the kernel receives lam AFTER the constraint; autograd carries the gradient back through the constraint
        return [dict(kind=self._kind, add_f0=bool(self.add_init_f0))], [self.transform_param().reshape(())], []


class InverseBoxCoxFlow(BoxCoxFlow):
    """fk = sgn(lam*f0+1)|lam*f0+1|^(1/lam) [+ f0] (reference flow.py:423-446)."""

    _kind = 'invboxcox'

    def __init__(self, init_lam, add_init_f0, constraint=None):
        super().__init__(init_lam, add_init_f0, constraint)

    def forward(self, f0, X=None):
        lam = self.transform_param()
        aux = lam * f0 + 1
        sgn = torch.sign(aux)
        fk = sgn * torch.pow(sgn * aux, 1. / lam)
        return fk + f0 if self.add_init_f0 else fk


def _mlp(input_dim, cfg, n_out_nets):
    """`n_out_nets` independent MLPs X -> R as the reference builds them (flow.py:853-871): num_H hidden layers of
    apply_linear(in, H, act, drop=DR, bn=BN) followed by a linear read-out to one unit."""
    BN, DR = cfg.get('batch_norm', 0), cfg.get('dropout', 0.0)
    H, act = cfg.get('hidden_dim', input_dim), cfg.get('hidden_activation', 'relu')
    num_H = cfg.get('num_hidden_layers', 1)
    inference = cfg.get('inference', 'MC_dropout')
    if inference != 'MC_dropout':
        raise NotImplementedError("only 'MC_dropout' input-dependent flows are in scope (every shipped configuration, "
                                  "exp_config.py:13,26); got %r" % (inference,))
    nets = []
    for _ in range(n_out_nets):
        layers, d = [], input_dim
        for _h in range(num_H):
            layers.append(apply_linear(d, H, act, shape=None, std=0.0, drop=DR, bn=BN))
            d = H
        layers.append(apply_linear(H, 1, 'linear', shape=None, std=0.0, drop=0.0, bn=0))
        nets.append(nn.Sequential(*layers))
    return nets, inference, BN == 1


class _RowParamMixin:
    """Shared plumbing of the input-dependent variants: scalar parameters serve only the initialiser and are detached
    from the model by `turn_off_initializer_parameters` (reference flow.py:924-934)."""

    _names = ()

    def turn_off_initializer_parameters(self):
        if self.input_dependent and not self.parameters_are_turn_off:
            for n in self._names:
                setattr(self, n + '_untracked', getattr(self, n).data.detach())
                setattr(self, n, None)
            self.parameters_are_turn_off = True

    def _nets(self):
        return [getattr(self, 'NNets_' + n) for n in self._names]

    dropout_masks = None          # parity mode: explicit keep-mask (n_nets, n_hidden_layers, rows, hidden) for the next calls
    last_dropout_masks = None     # the keep-mask the last fused evaluation used (exported by the kernel)

    def _net_outputs(self, X):
        """theta(x_n) of this layer's MLPs (reference flow.py:949-950), one tensor (rows,) per parameter.

        CUDA inputs run the fused kernel (tgp_flow_mlp_forward / _backward: all nets of the layer in one launch, MC-dropout
        from a Philox stream with the mask exported to `last_dropout_masks`, or from `dropout_masks` when set — the
        reference's masks come from torch's global RNG, so bit-parity needs them passed in; analytic backward).  Anything
        the kernel does not cover (batch-norm nets, CPU tensors used by the initialisers) runs the torch modules."""
        from ... import functional as Fn
        nets = self._nets()
        if X.is_cuda and X.dim() == 2 and cg.flow_mlp_in_kernel:
            spec = Fn.mlp_spec(nets)
            if spec is not None:
                out = Fn.flow_mlp(self, nets, spec, X, self.dropout_masks)
                return [out[:, i] for i in range(len(nets))]
        # under row sharding each rank evaluates the MLPs on its own rows: their parameter gradients are summed over
        # ranks in the backward (functional.synced_module_call); single process: a plain module call
        return [Fn.synced_module_call(net, X).squeeze(dim=-1) for net in nets]

    def forward_initializer(self, X):
        if not self.input_dependent:
            return 0.0
        return sum(((out - getattr(self, n).detach()) ** 2).mean()
                   for n, out in zip(self._names, (net(X) for net in self._nets())))

    def KLD(self):
        return 0.0


class TanhFlow(_RowParamMixin, Flow):
    """fk = a + b*tanh((f0-c)/d) [+ f0]; b, d -> softplus under set_restrictions."""
    _names = ('a', 'b', 'c', 'd')

    def __init__(self, init_a, init_b, init_c, init_d, add_init_f0, set_restrictions, input_dependent=False,
                 input_dim=-1, input_dependent_config={}):
        super().__init__()
        if input_dependent:
            assert input_dim > 0, 'Set input dimension if input_dependent = True'
            nets, self.inference, self.is_using_bn = _mlp(input_dim, input_dependent_config, 4)
            self.NNets_a, self.NNets_b, self.NNets_c, self.NNets_d = nets
            self.parameters_are_turn_off = False
        for n, val in zip(self._names, (init_a, init_b, init_c, init_d)):
            setattr(self, n, nn.Parameter(torch.tensor(val, dtype=cg.dtype)))
        self.set_restrictions = True if add_init_f0 else set_restrictions
        self.add_init_f0 = add_init_f0
        self.input_dependent = input_dependent

    def _params(self, X):
        if self.input_dependent:
            assert X is not None, 'Set X to value'
            assert self.parameters_are_turn_off, 'Call turn_off_initializer_parameters before using the flow'
            a, b, c, d = self._net_outputs(X)
        else:
            a, b, c, d = self.a, self.b, self.c, self.d
        if self.set_restrictions:
            b, d = softplus(b), softplus(d)
        return a, b, c, d

    def forward(self, f0, X=None):
        a, b, c, d = self._params(X)
        fk = a + b * torch.tanh((f0 - c) / d)
        return fk + f0 if self.add_init_f0 else fk

    def describe(self, X=None, n_mc=1):
        if not self.set_restrictions:
            raise NotImplementedError('the fused tanh layer assumes set_restrictions=True (StepTanhL always sets it)')
        lay = dict(kind='tanh_step', n_steps=1, add_f0=bool(self.add_init_f0), per_row=bool(self.input_dependent))
        if self.input_dependent:
            return [lay], [], self._net_outputs(X)
        return [lay], [self.a, self.b, self.c, self.d], []


class Sinh_ArcsinhFlow(_RowParamMixin, Flow):
    """fk = sinh(b*asinh(f0) - a) [+ f0]; b -> softplus(b) under set_restrictions; a, b optionally MLPs of X."""
    _names = ('a', 'b')

    def __init__(self, init_a, init_b, add_init_f0, set_restrictions, input_dependent=False, input_dim=-1,
                 input_dependent_config={}):
        super().__init__()
        if input_dependent:
            assert input_dim > 0, 'Set input dimension if input_dependent = True'
            nets, self.inference, self.is_using_bn = _mlp(input_dim, input_dependent_config, 2)
            self.NNets_a, self.NNets_b = nets
            self.parameters_are_turn_off = False
        self.a = nn.Parameter(torch.tensor(init_a, dtype=cg.dtype))
        self.b = nn.Parameter(torch.tensor(init_b, dtype=cg.dtype))
        self.set_restrictions = True if add_init_f0 else set_restrictions
        self.add_init_f0 = add_init_f0
        self.input_dependent = input_dependent

    def asinh(self, f):
        return torch.log(f + (f ** 2 + 1) ** 0.5)      # the reference's formulation (flow.py:904-905)

    def _params(self, X):
        if self.input_dependent:
            assert X is not None, 'Set X to value'
            assert self.parameters_are_turn_off, 'Call turn_off_initializer_parameters before using the flow'
            if self.is_using_bn:
                lead = X.shape[:-1]
                a, b = (o.view(lead) for o in self._net_outputs(X.reshape(-1, X.shape[-1])))
            else:
                a, b = self._net_outputs(X)
        else:
            a, b = self.a, self.b
        return a, (softplus(b) if self.set_restrictions else b)

    def forward(self, f0, X=None):
        a, b = self._params(X)
        fk = torch.sinh(b * self.asinh(f0) - a)
        return fk + f0 if self.add_init_f0 else fk

    def inverse(self, f):
        b = softplus(self.b) if self.set_restrictions else self.b
        return torch.sinh(1 / b * (self.asinh(f) + self.a))

    def describe(self, X=None, n_mc=1):
        lay = dict(kind='sal', restrict=bool(self.set_restrictions), add_f0=bool(self.add_init_f0),
                   per_row=bool(self.input_dependent))
        if self.input_dependent:
            assert self.parameters_are_turn_off, 'Call turn_off_initializer_parameters before using the flow'
            return [lay], [], self._net_outputs(X)
        return [lay], [self.a, self.b], []


class switch_off(nn.Module):
    """Per-step (scale, bias) of a StepFlow: trainable only for steps without their own scale (flow.py:1130-1149)."""

    def __init__(self, is_trainable, n_steps):
        super().__init__()
        self.is_trainable = is_trainable
        if is_trainable:
            self.a = nn.Parameter(inv_softplus(torch.tensor(1.0 / float(n_steps), dtype=cg.dtype)))
            self.b = nn.Parameter(torch.tensor(0.0, dtype=cg.dtype))

    def forward(self):
        if self.is_trainable:
            return softplus(self.a), self.b
        return 1.0, 0.0


class StepFlow(Flow):
    """fk = sum_i [s_i * g_i(f0) + t_i] (+ f0): linear combination of elementary flows (flow.py:1039-1128)."""

    def __init__(self, flow_arr, add_init_f0):
        super().__init__()
        assert isinstance(add_init_f0, bool), 'add_init_f0 must be boolean'
        self.add_init_f0 = add_init_f0
        self.switch_off = nn.ModuleList()
        n_steps = len(flow_arr)
        for step in flow_arr:
            if isinstance(step, (list, tuple)):
                name, params = step
                restricted = params.get('set_restrictions', False)
            else:
                name = {Sinh_ArcsinhFlow: 'sinh_arcsinh', TanhFlow: 'tanh', BoxCoxFlow: 'boxcox'}[type(step)]
                restricted = getattr(step, 'set_restrictions', False)
            assert name != 'step_flow', 'cannot combine step flow with step flow'
            assert name in ('boxcox', 'inverseboxcox') or restricted, \
                'set_restrictions must be True. Got false for flow {}'.format(name)
            # flows without their own scale and bias get a trainable (scale, bias) so that they can be switched off
            self.switch_off.append(switch_off(name in ('boxcox', 'sinh_arcsinh', 'inverseboxcox'), n_steps))
        if isinstance(flow_arr[0], (list, tuple)):
            self.flow_arr = nn.ModuleList(instance_flow(flow_arr, is_composite=False))
        else:
            self.flow_arr = nn.ModuleList(flow_arr)

    def forward(self, f0, X=None):
        fk = 0.0
        for sw, fl in zip(self.switch_off, self.flow_arr):
            s, t = sw()
            fk = fk + (s * fl.forward(f0, X=X) + t)
        return fk + f0 if self.add_init_f0 else fk

    def forward_initializer(self, X):
        return sum(fl.forward_initializer(X) for fl in self.flow_arr)

    @property
    def input_dependent(self):
        return None

    @input_dependent.setter
    def input_dependent(self, value):
        for fl in self.flow_arr:
            fl.input_dependent = value

    def turn_off_initializer_parameters(self):
        for fl in self.flow_arr:
            fl.turn_off_initializer_parameters()

    def describe(self, X=None, n_mc=1):
        steps = list(self.flow_arr)
        if not all(isinstance(s, TanhFlow) and not s.add_init_f0 and s.set_restrictions for s in steps):
            return self._describe_group(X, n_mc)
        per_row = [bool(s.input_dependent) for s in steps]
        if any(per_row) and not all(per_row):
            raise NotImplementedError('mixed input-dependent / global tanh steps')
        lay = dict(kind='tanh_step', n_steps=len(steps), add_f0=bool(self.add_init_f0), per_row=per_row[0])
        if per_row[0]:
            rows = []
            for s in steps:
                rows += s._net_outputs(X)
            return [lay], [], rows
        glob = []
        for s in steps:
            glob += [s.a, s.b, s.c, s.d]
        return [lay], glob, []

    def _describe_group(self, X, n_mc):
        """General step flow (StepSAL / StepArcSL / StepBoxCoxL / StepInverseBoxCoxL / StepAllL, flows.py:284-492): a group
        header followed by one layer per member; members with a trainable switch_off carry its [scale, bias]."""
        layers = [dict(kind='step_group', n_steps=len(self.flow_arr), add_f0=bool(self.add_init_f0))]
        glob, rows = [], []
        for sw, fl in zip(self.switch_off, self.flow_arr):
            if isinstance(fl, (AffineFlow, IdentityFlow, StepFlow)):
                raise NotImplementedError('%s as a step member' % type(fl).__name__)
            l, g, r = fl.describe(X, n_mc)
            assert len(l) == 1
            if l[0]['kind'] == 'tanh_step' or isinstance(fl, TanhFlow):
                l[0] = dict(l[0], kind='tanh_step', n_steps=1)
            if sw.is_trainable:
                if l[0].get('per_row', False):
                    raise NotImplementedError('input-dependent step members with a trainable switch_off')
                l[0] = dict(l[0], switch=True)
                g = list(g) + [sw.a, sw.b]
            layers += l
            glob += g
            rows += r
        return layers, glob, rows
