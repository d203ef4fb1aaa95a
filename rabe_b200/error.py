"""RabeError mirror (/root/reference/src/error.rs:10-80): a details string, nothing else."""


class RabeError(Exception):
    def __init__(self, details):
        super().__init__(details)
        self.details = details
