"""Functional-network surface (level B1 of the drop-in boundary).

The reference's ``VGGReLUNormNetwork`` (``meta_neural_network_architectures.py:545-688``) is a
4-block conv3x3 -> BatchNorm(batch statistics, per-step gamma/beta) -> leaky-ReLU -> maxpool2 net
followed by a linear layer, all taking externally supplied ("fast") weights.  Here the classes keep
the reference's module tree and parameter names -- so ``state_dict`` keys, shapes, registration
order (= Adam parameter order) and initialisation RNG consumption are identical -- but they are
parameter containers: the arithmetic of the path runs in the CUDA engine (``csrc/``), which sees
these parameters as one flat buffer.

Initialisation restates reference ``:62-66`` (xavier_uniform_ conv weight, zero bias),
``:114-118`` (xavier_uniform_ linear weights), ``:177-192`` (running_mean zeros; running_var ones
per-step but ZEROS in shared mode; beta zeros; gamma ones).
"""
import torch
import torch.nn as nn


class MetaConv2dLayer(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride, padding, use_bias, groups=1, dilation_rate=1):
        super().__init__()
        if int(kernel_size) != 3 or int(stride) != 1 or int(padding) != 1 or int(groups) != 1 or int(dilation_rate) != 1:
            raise NotImplementedError("the B200 engine implements the shipped configuration only: 3x3, stride 1, "
                                      "padding 1 (max_pooling=true, conv_padding=true)")
        self.stride, self.padding, self.dilation_rate, self.groups, self.use_bias = 1, 1, 1, 1, bool(use_bias)
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels, 3, 3))
        nn.init.xavier_uniform_(self.weight)
        if self.use_bias:
            self.bias = nn.Parameter(torch.zeros(out_channels))


class MetaLinearLayer(nn.Module):
    def __init__(self, input_shape, num_filters, use_bias):
        super().__init__()
        _, c = input_shape
        self.use_bias = bool(use_bias)
        self.weights = nn.Parameter(torch.ones(num_filters, int(c)))
        nn.init.xavier_uniform_(self.weights)
        if self.use_bias:
            self.bias = nn.Parameter(torch.zeros(num_filters))


class MetaBatchNormLayer(nn.Module):
    def __init__(self, num_features, device, args, eps=1e-5, momentum=0.1, affine=True, track_running_stats=True,
                 meta_batch_norm=True, no_learnable_params=False, use_per_step_bn_statistics=False):
        super().__init__()
        if getattr(args, "enable_inner_loop_optimizable_bn_params", False):
            raise NotImplementedError("enable_inner_loop_optimizable_bn_params is outside the accelerated path "
                                      "(no shipped config sets it)")
        self.num_features, self.eps, self.momentum = int(num_features), eps, momentum
        self.use_per_step_bn_statistics = bool(use_per_step_bn_statistics)
        S = int(args.number_of_training_steps_per_iter)
        shape = (S, self.num_features) if self.use_per_step_bn_statistics else (self.num_features,)
        self.running_mean = nn.Parameter(torch.zeros(shape), requires_grad=False)
        self.running_var = nn.Parameter(torch.ones(shape) if self.use_per_step_bn_statistics else torch.zeros(shape),
                                        requires_grad=False)
        self.bias = nn.Parameter(torch.zeros(shape), requires_grad=bool(args.learnable_bn_beta))
        self.weight = nn.Parameter(torch.ones(shape), requires_grad=bool(args.learnable_bn_gamma))
        if self.use_per_step_bn_statistics:
            # Reference quirk: while building itself the reference network pushes an all-zero dummy batch
            # through every block twice with num_step=0 (build_block :365 and build_network :603), and
            # F.batch_norm's EMA side effect leaves running_var[0] = 0.9 * 0.9 * 1 (fp32) in a fresh model.
            with torch.no_grad():
                for _ in range(2):
                    self.running_var[0].mul_(1.0 - momentum)


class MetaConvNormLayerReLU(nn.Module):
    def __init__(self, input_shape, num_filters, kernel_size, stride, padding, use_bias, args, normalization=True,
                 meta_layer=True, no_bn_learnable_params=False, device=None):
        super().__init__()
        if not normalization or getattr(args, "norm_layer", "batch_norm") != "batch_norm":
            raise NotImplementedError("only norm_layer='batch_norm' is on the accelerated path")
        self.layer_dict = nn.ModuleDict()
        self.conv = MetaConv2dLayer(in_channels=int(input_shape[1]), out_channels=num_filters, kernel_size=kernel_size,
                                    stride=stride, padding=padding, use_bias=use_bias)
        self.norm_layer = MetaBatchNormLayer(num_filters, device=device, args=args,
                                             use_per_step_bn_statistics=args.per_step_bn_statistics)


class VGGReLUNormNetwork(nn.Module):
    def __init__(self, im_shape, num_output_classes, args, device, meta_classifier=True):
        super().__init__()
        if not args.max_pooling:
            raise NotImplementedError("strided-conv / avg-pool variant (max_pooling=false) is outside the "
                                      "accelerated path (all shipped configs use max pooling)")
        _, c, h, w = im_shape
        self.args, self.device = args, device
        self.num_stages = int(args.num_stages)
        self.cnn_filters = int(args.cnn_num_filters)
        self.num_output_classes = int(num_output_classes)
        self.layer_dict = nn.ModuleDict()
        shape = [int(im_shape[0]), int(c), int(h), int(w)]
        for i in range(self.num_stages):
            self.layer_dict["conv%d" % i] = MetaConvNormLayerReLU(
                input_shape=shape, num_filters=self.cnn_filters, kernel_size=3, stride=1,
                padding=int(bool(args.conv_padding)), use_bias=True, args=args, device=device)
            shape = [shape[0], self.cnn_filters, shape[2] // 2, shape[3] // 2]
            if shape[2] < 1 or shape[3] < 1:
                raise ValueError("image too small for %d stages" % self.num_stages)
        self.encoder_features_shape = list(shape)
        feat = shape[1] * shape[2] * shape[3]
        self.layer_dict["linear"] = MetaLinearLayer(input_shape=(shape[0], feat), num_filters=self.num_output_classes,
                                                    use_bias=True)

    def forward(self, x, num_step, params=None, training=False, backup_running_statistics=False):
        raise NotImplementedError(
            "VGGReLUNormNetwork is a parameter container here: the conv/BN/leaky-ReLU/maxpool/linear arithmetic "
            "of this network runs inside the CUDA engine, driven by MAMLFewShotClassifier.run_train_iter / "
            "run_validation_iter (see INTEGRATION.md)")

    def zero_grad(self, params=None):
        if params is None:
            for p in self.parameters():
                p.grad = None

    def restore_backup_stats(self):
        """Evaluation never commits running statistics in the engine, so there is nothing to restore."""
        return None
