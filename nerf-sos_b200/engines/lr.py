"""Drop-in for engines/lr.py:LRScheduler (reference :4-23): lr_g = init_lr_g * decay_rate ** (step / decay_steps)."""


class LRScheduler:
    def __init__(self, optimizer, init_lr, decay_rate, decay_steps):
        self.init_lr = [init_lr] if isinstance(init_lr, (int, float)) else list(init_lr)
        self.decay_rate, self.decay_steps, self.optimizer = decay_rate, decay_steps, optimizer
        assert len(optimizer.param_groups) == len(self.init_lr), "Number of lr does not match number of param groups."

    def step(self, step):
        for lr, group in zip(self.init_lr, self.optimizer.param_groups):
            group["lr"] = lr * (self.decay_rate ** (step / self.decay_steps))
