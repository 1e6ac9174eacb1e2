"""Flow *specification* generators — lists of `(name, init_dict)` consumed by `instance_flow`
(reference code/dsp/flows.py: common_config :11-32, set_input_dependent_config :34-69, build_chain :71-109, SAL :115-136,
BoxCoxL / InverseBoxCoxL / ArcSL :140-214, StepTanhL :239-277, StepSAL / StepArcSL / StepBoxCoxL / StepInverseBoxCoxL /
StepAllL :284-491).  Draws from numpy's global RNG in the same order as the reference so that a seeded run
starts from the same flow parameters.
"""
import numpy
import torch

from .models.flow import *  # noqa: F401,F403  (the reference re-exports the flow classes from here)
from .utils import inv_softplus

_ID_KEYS = ('batch_norm', 'dropout', 'hidden_dim', 'hidden_activation', 'num_hidden_layers', 'inference')


def common_config(options):
    return (options.get('set_res', False), options.get('add_f0', False), options.get('init_random', False),
            options.get('constraint', None))


def set_input_dependent_config(options):
    input_dependent = bool(options.get('input_dependent', False))
    if input_dependent:
        assert 'input_dim' in options, 'You set to use input_dependent flows but the input dimension is not provided.'
    cfg = {k: options[k] for k in _ID_KEYS if k in options}
    return input_dependent, options.get('input_dim', -1), cfg


def SAL(num_blocks, **kwargs):
    """num_blocks x [sinh_arcsinh, affine]; the default initialisation is the identity map."""
    set_res, addf0, init_random, _ = common_config(kwargs)
    input_dependent, input_dim, id_cfg = set_input_dependent_config(kwargs)
    blocks = []
    for _ in range(num_blocks):
        if init_random:
            a_aff, b_aff = numpy.random.randn(2)
            a_sal, b_sal = numpy.random.randn(2)
        else:
            a_aff, b_aff, a_sal, b_sal = 1.0, 0.0, 0.0, 1.0
        blocks.append(('sinh_arcsinh', {'init_a': a_sal, 'init_b': b_sal, 'add_init_f0': addf0,
                                        'set_restrictions': set_res, 'input_dependent': input_dependent,
                                        'input_dim': input_dim, 'input_dependent_config': id_cfg}))
        blocks.append(('affine', {'init_a': a_aff, 'init_b': b_aff, 'set_restrictions': set_res}))
    return blocks


def StepTanhL(num_blocks, num_steps, **kwargs):
    """num_blocks x [step_flow(num_steps tanh terms), affine]; tanh steps are always restricted (b, d > 0)."""
    _, addf0, init_random, _ = common_config(kwargs)
    if 'set_res' in kwargs:
        assert kwargs['set_res'] is True, 'In the step tanh flow set_res has to be True for num_steps > 1'
    input_dependent, input_dim, id_cfg = set_input_dependent_config(kwargs)
    blocks = []
    for _ in range(num_blocks):
        steps = []
        for _s in range(num_steps):
            e1, e2, e3, e4 = numpy.multiply(numpy.random.randn(4, ), numpy.array([1.0, 1.0, 1.0, 1.0]))
            if not init_random:
                e2 = inv_softplus(torch.abs(torch.tensor((e2 + 1.0) / float(num_steps)))).item()
                e4 = inv_softplus(torch.abs(torch.tensor((e4 + 1.0) / float(num_steps)))).item()
            steps.append(('tanh', {'init_a': e1, 'init_b': e2, 'init_c': e3, 'init_d': e4, 'add_init_f0': False,
                                   'set_restrictions': True, 'input_dependent': input_dependent,
                                   'input_dim': input_dim, 'input_dependent_config': id_cfg}))
        a_aff, b_aff = numpy.random.randn(2) if init_random else (1.0, 0.0)
        blocks.append(('step_flow', {'flow_arr': steps, 'add_init_f0': addf0}))
        blocks.append(('affine', {'init_a': a_aff, 'init_b': b_aff, 'set_restrictions': False}))
    return blocks


def BoxCoxL(num_blocks, **kwargs):
    """num_blocks x [boxcox, affine] (reference flows.py:140-163)."""
    set_res, addf0, init_random, constraint = common_config(kwargs)
    blocks = []
    for _ in range(num_blocks):
        if init_random:
            a_aff, b_aff = numpy.random.randn(2)
            init_lam = numpy.random.randn(1) + 1.
            constraint = None
        else:
            a_aff, b_aff, init_lam = 1.0, 0.0, 5.0
        blocks.append(('boxcox', {'init_lam': init_lam, 'add_init_f0': addf0, 'constraint': constraint}))
        blocks.append(('affine', {'init_a': a_aff, 'init_b': b_aff, 'set_restrictions': set_res}))
    return blocks


def InverseBoxCoxL(num_blocks, **kwargs):
    """num_blocks x [inverseboxcox, affine] (reference flows.py:167-189)."""
    set_res, addf0, init_random, constraint = common_config(kwargs)
    blocks = []
    for _ in range(num_blocks):
        if init_random:
            a_aff, b_aff = numpy.random.randn(2)
            init_lam = numpy.random.randn(1) + 1.
        else:
            a_aff, b_aff, init_lam = 1.0, 0.0, 5.0
        blocks.append(('inverseboxcox', {'init_lam': init_lam, 'add_init_f0': addf0, 'constraint': constraint}))
        blocks.append(('affine', {'init_a': a_aff, 'init_b': b_aff, 'set_restrictions': set_res}))
    return blocks


def ArcSL(num_blocks, **kwargs):
    """num_blocks x [arcsinh, affine] (reference flows.py:193-214)."""
    set_res, addf0, init_random, _ = common_config(kwargs)
    blocks = []
    for _ in range(num_blocks):
        if init_random:
            a_aff, b_aff = numpy.random.randn(2)
            a_arc, b_arc, c_arc, d_arc = numpy.random.randn(4)
        else:
            a_aff, b_aff = 1.0, 0.0
            a_arc, b_arc, c_arc, d_arc = numpy.random.randn(4)
            b_arc += 1
            d_arc += 1
        blocks.append(('arcsinh', {'init_a': a_arc, 'init_b': b_arc, 'init_c': c_arc, 'init_d': d_arc, 'add_init_f0': addf0,
                                   'set_restrictions': set_res}))
        blocks.append(('affine', {'init_a': a_aff, 'init_b': b_aff, 'set_restrictions': set_res}))
    return blocks


def _step_blocks(num_blocks, kwargs, draw_members):
    """num_blocks x [step_flow(members), affine(unrestricted)] — the shape shared by every Step*L generator.  `draw_members`
    draws one block's member list; the affine draw follows it, as in the reference."""
    _, addf0, init_random, _ = common_config(kwargs)
    if 'set_res' in kwargs:
        assert kwargs['set_res'] is True, 'In the step flows set_res has to be True for num_steps > 1'
    blocks = []
    for _ in range(num_blocks):
        members = draw_members()
        a_aff, b_aff = numpy.random.randn(2) if init_random else (1.0, 0.0)
        blocks.append(('step_flow', {'flow_arr': members, 'add_init_f0': addf0}))
        blocks.append(('affine', {'init_a': a_aff, 'init_b': b_aff, 'set_restrictions': False}))
    return blocks


def _softplus_inverse_of_share(e, num_steps):
    return inv_softplus(torch.abs(torch.tensor((e + 1.0) / float(num_steps)))).item()


def _draw_abcd(init_random, num_steps):
    """Four N(0,1) draws; by default the scale-like ones (b, d) start at softplus^-1(|e + 1| / num_steps)."""
    e1, e2, e3, e4 = numpy.multiply(numpy.random.randn(4, ), numpy.array([1.0, 1.0, 1.0, 1.0]))
    if not init_random:
        e2, e4 = _softplus_inverse_of_share(e2, num_steps), _softplus_inverse_of_share(e4, num_steps)
    return {'init_a': e1, 'init_b': e2, 'init_c': e3, 'init_d': e4, 'add_init_f0': False, 'set_restrictions': True}


def _draw_sal(init_random):
    a_sal, b_sal = numpy.random.randn(2)
    if not init_random:
        b_sal = inv_softplus(torch.abs(torch.tensor(b_sal + 1.0))).item()
    return {'init_a': a_sal, 'init_b': b_sal, 'add_init_f0': False, 'set_restrictions': True}


def _draw_lam(init_random, addf0, constraint):
    init_lam = numpy.random.randn(1, )
    if not init_random:
        init_lam += 5.0
    return {'init_lam': init_lam, 'add_init_f0': addf0, 'constraint': constraint}      # members add f0 themselves too


def StepSAL(num_blocks, num_steps, **kwargs):
    """num_blocks x [step_flow(num_steps sinh_arcsinh, each behind a trainable switch_off), affine] (flows.py:284-316)."""
    init_random = common_config(kwargs)[2]
    return _step_blocks(num_blocks, kwargs, lambda: [('sinh_arcsinh', _draw_sal(init_random)) for _ in range(num_steps)])


def StepArcSL(num_blocks, num_steps, **kwargs):
    """sum_i [a_i + b_i asinh((f - c_i) / d_i)], then affine (flows.py:322-354)."""
    init_random = common_config(kwargs)[2]
    return _step_blocks(num_blocks, kwargs,
                        lambda: [('arcsinh', _draw_abcd(init_random, num_steps)) for _ in range(num_steps)])


def StepBoxCoxL(num_blocks, num_steps, **kwargs):
    """Sum of Box-Cox members behind trainable switch_offs, then affine (flows.py:358-388)."""
    _, addf0, init_random, constraint = common_config(kwargs)
    return _step_blocks(num_blocks, kwargs,
                        lambda: [('boxcox', _draw_lam(init_random, addf0, constraint)) for _ in range(num_steps)])


def StepInverseBoxCoxL(num_blocks, num_steps, **kwargs):
    """Sum of inverse Box-Cox members behind trainable switch_offs, then affine (flows.py:391-421)."""
    _, addf0, init_random, constraint = common_config(kwargs)
    return _step_blocks(num_blocks, kwargs,
                        lambda: [('inverseboxcox', _draw_lam(init_random, addf0, constraint)) for _ in range(num_steps)])


def StepAllL(num_blocks, **kwargs):
    """One member of each family — inverse Box-Cox, Box-Cox, arcsinh, sinh-arcsinh, tanh — then affine (flows.py:425-491).
    The reference returns from inside its block loop, i.e. always ONE block whatever num_blocks is; kept."""
    _, addf0, init_random, constraint = common_config(kwargs)

    def members():
        return [('inverseboxcox', _draw_lam(init_random, addf0, constraint)),
                ('boxcox', _draw_lam(init_random, addf0, constraint)),
                ('arcsinh', _draw_abcd(init_random, 5)),
                ('sinh_arcsinh', _draw_sal(init_random)),
                ('tanh', _draw_abcd(init_random, 5))]
    return _step_blocks(min(num_blocks, 1), kwargs, members)


def build_chain(flow_combination, num_blocks, **kwargs):
    """Combined chains of the reference's launch scripts (reference flows.py:71-109)."""
    blocks = []
    for _ in range(num_blocks):
        if flow_combination == 'SAL_BCL':
            blocks.extend(SAL(1)); blocks.extend(BoxCoxL(1, constraint=kwargs['constraint']))          # noqa: E702
        elif flow_combination == 'SAL_InvBCL':
            blocks.extend(SAL(1)); blocks.extend(InverseBoxCoxL(1, constraint=kwargs['constraint']))   # noqa: E702
        elif flow_combination == 'SAL_AL':
            blocks.extend(SAL(1)); blocks.extend(ArcSL(1))                                             # noqa: E702
        elif flow_combination == 'BCL_AL':
            blocks.extend(BoxCoxL(1, constraint=kwargs['constraint'])); blocks.extend(ArcSL(1))        # noqa: E702
        elif flow_combination == 'InvBCL_AL':
            blocks.extend(InverseBoxCoxL(1, constraint=kwargs['constraint'])); blocks.extend(ArcSL(1))  # noqa: E702
    return blocks
