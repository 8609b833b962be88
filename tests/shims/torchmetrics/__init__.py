"""torchmetrics.Metric is a base class of evaluators the evaluation path constructs but never calls (TEST INFRASTRUCTURE)."""


class Metric:
    def __init__(self, *a, **k):
        pass

    def add_state(self, *a, **k):
        pass
