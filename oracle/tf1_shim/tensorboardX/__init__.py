"""Stand-in for tensorboardX (absent from this image) so that the reference's launch scripts can be parsed/imported
by the oracle tooling -- TEST INFRASTRUCTURE ONLY."""


class SummaryWriter:
    def __init__(self, *a, **k):
        self.scalars = []

    def add_scalar(self, tag, value, step=None):
        self.scalars.append((tag, float(value), step))
