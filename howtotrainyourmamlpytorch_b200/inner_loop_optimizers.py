"""Inner-loop optimiser surface (level B2 of the drop-in boundary).

``LSLRGradientDescentLearningRule`` keeps the reference's constructor, ``initialise`` and
``update_params`` (reference ``inner_loop_optimizers.py:55-113``): one learnable learning-rate
vector of length ``S+1`` per adaptable tensor, ``theta' = theta - alpha[name][step] * g``.

On the hot path this arithmetic does not run here: the engine fuses it into the kernel that reduces
the weight-gradient partials (``csrc/kernels_param.cu: param_reduce_kernel``, mode PR_UPDATE).  The
module below owns the ``alpha`` parameters (same ``state_dict`` names as the reference, stored as
views of the system's flat meta-parameter buffer) and offers ``update_params`` as the same pure
function for callers that use the operator directly.
"""
import torch
import torch.nn as nn


class LSLRGradientDescentLearningRule(nn.Module):
    def __init__(self, device, total_num_inner_loop_steps, use_learnable_learning_rates, init_learning_rate=1e-3):
        super().__init__()
        assert init_learning_rate > 0.0, "learning_rate should be positive."
        self.device = device
        self.init_learning_rate = float(init_learning_rate)
        self.total_num_inner_loop_steps = int(total_num_inner_loop_steps)
        self.use_learnable_learning_rates = bool(use_learnable_learning_rates)
        self.names_learning_rates_dict = nn.ParameterDict()

    def initialise(self, names_weights_dict):
        """One ``[S+1]`` vector per inner-loop tensor, keyed by the tensor name with '.' -> '-'."""
        self.names_learning_rates_dict = nn.ParameterDict()
        for key in names_weights_dict.keys():
            self.names_learning_rates_dict[key.replace(".", "-")] = nn.Parameter(
                torch.full((self.total_num_inner_loop_steps + 1,), self.init_learning_rate),
                requires_grad=self.use_learnable_learning_rates)

    def reset(self):
        pass

    def update_params(self, names_weights_dict, names_grads_wrt_params_dict, num_step, tau=0.1):
        """Pure function: returns the dict of updated fast weights for inner step ``num_step``."""
        out = {}
        for key, grad in names_grads_wrt_params_dict.items():
            lr = self.names_learning_rates_dict[key.replace(".", "-")][num_step]
            out[key] = names_weights_dict[key] - lr * grad
        return out
