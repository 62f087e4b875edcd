"""stand-in for torch_ema: ExponentialMovingAverage with the package's semantics (decay warm-up by update count,
shadow parameters, store / copy_to / restore through `param.data.copy_`, state_dict round trip)"""
import torch


class ExponentialMovingAverage:
    def __init__(self, parameters, decay, use_num_updates=True):
        if decay < 0.0 or decay > 1.0:
            raise ValueError("Decay must be between 0 and 1")
        self.decay = decay
        self.num_updates = 0 if use_num_updates else None
        self._params = [p for p in parameters]
        self.shadow_params = [p.clone().detach() for p in self._params]
        self.collected_params = None

    def _get(self, parameters):
        return self._params if parameters is None else list(parameters)

    def update(self, parameters=None):
        params = self._get(parameters)
        decay = self.decay
        if self.num_updates is not None:
            self.num_updates += 1
            decay = min(decay, (1 + self.num_updates) / (10 + self.num_updates))
        one_minus = 1.0 - decay
        with torch.no_grad():
            for s, p in zip(self.shadow_params, params):
                if p.requires_grad:
                    s.sub_(one_minus * (s - p))

    def copy_to(self, parameters=None):
        for s, p in zip(self.shadow_params, self._get(parameters)):
            if p.requires_grad:
                p.data.copy_(s.data)

    def store(self, parameters=None):
        self.collected_params = [p.clone() for p in self._get(parameters)]

    def restore(self, parameters=None):
        if self.collected_params is None:
            raise RuntimeError("restore() called before store()")
        for c, p in zip(self.collected_params, self._get(parameters)):
            p.data.copy_(c.data)

    def state_dict(self):
        return {"decay": self.decay, "num_updates": self.num_updates, "shadow_params": self.shadow_params,
                "collected_params": self.collected_params}

    def load_state_dict(self, state):
        self.decay, self.num_updates = state["decay"], state["num_updates"]
        self.shadow_params = [s.to(p.device) for s, p in zip(state["shadow_params"], self._params)]
        self.collected_params = state.get("collected_params")
