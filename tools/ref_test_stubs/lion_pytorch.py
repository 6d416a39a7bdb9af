"""Stand-in for the third-party `lion_pytorch` package, which the reference's tests/test_optim.py imports as its
Lion baseline and which is not installed in this image (no network).  The published algorithm (Chen et al. 2023,
"Symbolic Discovery of Optimization Algorithms"), eager PyTorch:

    p <- p * (1 - lr * wd);  p <- p - lr * sign(beta1 * m + (1 - beta1) * g);  m <- beta2 * m + (1 - beta2) * g

Only tools/run_reference_tests.sh puts this directory on PYTHONPATH."""
import torch


class Lion(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-4, betas=(0.9, 0.99), weight_decay=0.0, **_unused):
        if lr <= 0.0:
            raise ValueError(f"Invalid learning rate: {lr}")
        super().__init__(params, dict(lr=lr, betas=betas, weight_decay=weight_decay))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            lr, (beta1, beta2), wd = group["lr"], group["betas"], group["weight_decay"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                state = self.state[p]
                if "exp_avg" not in state:
                    state["exp_avg"] = torch.zeros_like(p)
                m = state["exp_avg"]
                p.mul_(1 - lr * wd)
                p.add_(torch.sign(m * beta1 + p.grad * (1 - beta1)), alpha=-lr)
                m.mul_(beta2).add_(p.grad, alpha=1 - beta2)
        return loss
