"""Drop-in `network_arch` module (same class names, attributes and pickle layout as the
reference's src/network_arch.py) so that checkpoints written here load in the reference's
utils/eval_nn_on_grid.py / eval_nn_on_sample_data.py and vice versa.

The modules are plain parameter containers on the host (fp64, as the reference keeps them,
src/pytorch_eigen_solver.py:374-376); the training step itself never calls `forward` -- it runs
in libeigenpde_b200.so.  `forward` is kept for evaluation scripts and small CPU checks.
"""
import torch


class MySequential(torch.nn.Module):
    """One fully connected net  Linear -> act -> ... -> Linear  (reference: network_arch.py:4-69)."""

    def __init__(self, size_list, activation_name):
        super().__init__()
        self.nn_dims = size_list
        n_lin = len(size_list) - 1
        self.linears = torch.nn.ModuleList(torch.nn.Linear(size_list[j], size_list[j + 1]) for j in range(n_lin))
        act_cls = getattr(torch.nn, activation_name, None)
        if act_cls is None:
            print('[Warning] No activation function with name %s. Use CELU instead' % activation_name)
            act_cls = torch.nn.CELU
        else:
            print('[Info] Build NN with activation function %s.' % activation_name)
        self.activations = torch.nn.ModuleList(act_cls() for _ in range(n_lin - 1))
        # N(0, 0.5) for weights and biases; the default Linear init above has already consumed
        # its share of the torch RNG stream, exactly as in the reference (:25, :37-39)
        for lin in self.linears:
            torch.nn.init.normal_(lin.weight, 0, 0.5)
            torch.nn.init.normal_(lin.bias, 0, 0.5)

    def shift_and_normalize(self, mean, var):
        """f <- (f - mean) / sqrt(var), by editing the last layer (reference :41-50)."""
        with torch.no_grad():
            last = self.linears[-1]
            scale = torch.sqrt(torch.as_tensor(var, dtype=last.weight.dtype))
            last.bias.sub_(mean)
            last.weight.div_(scale)
            last.bias.div_(scale)

    def forward(self, x):
        h = x.pre_processing_layer()
        last = len(self.linears) - 1
        for j, lin in enumerate(self.linears):
            h = lin(h)
            if j < last:
                h = self.activations[j](h)
        return h


class MyNet(torch.nn.Module):
    """d_out independent MySequential nets, outputs concatenated (reference: network_arch.py:71-117)."""

    def __init__(self, size_list, activation_name, d_out):
        super().__init__()
        self.d_out = d_out
        self.nets = torch.nn.ModuleList(MySequential(size_list, activation_name) for _ in range(d_out))

    def shift_and_normalize(self, mean_list, var_list):
        for net, m, v in zip(self.nets, mean_list, var_list):
            net.shift_and_normalize(m, v)

    def forward(self, x):
        return torch.cat([net(x) for net in self.nets], dim=1)
