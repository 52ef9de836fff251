"""``SAC_file/SAC_add_discrete.py`` — SAC with an added hand-made discrete variant.

For continuous action spaces the file's class is ``SAC_file/SAC.py``'s, statement for statement: running the unmodified reference
class on the seeds of ``oracle/make_golden.py::gen_sac`` reproduces ``tests/golden/sac.npz`` bit for bit
(``tests/test_launcher.py::test_sac_add_discrete_continuous_is_sac``), so the continuous case IS :class:`freerl_b200.SAC.SAC`
on the fused actor-critic kernel.  The discrete ``hands_on`` variant (softmax actor, per-action twin V heads,
``Q = sum(probs * min(V1, V2))``, ``SAC_add_discrete.py:137-177,299-341``) is not implemented: the constructor raises.
"""
from .SAC import SAC as _SAC


class SAC(_SAC):
    def __init__(self, dim_info, is_continue, actor_lr, critic_lr, buffer_size, device, trick=None, mode=None):
        if not is_continue:
            raise NotImplementedError("SAC_add_discrete: the discrete 'hands_on' variant is not implemented on the fused kernel "
                                      "(continuous action spaces run freerl_b200.SAC.SAC)")
        self.discrete_type = {'hands_on': True, 'other': False}
        super().__init__(dim_info, is_continue, actor_lr, critic_lr, buffer_size, device, trick=trick, mode=mode)

    @staticmethod
    def load(dim_info, is_continue, model_dir, trick=None, device=None):
        if not is_continue:
            raise NotImplementedError("SAC_add_discrete: the discrete 'hands_on' variant is not implemented")
        return _SAC.load(dim_info, is_continue, model_dir, trick=trick, device=device)
