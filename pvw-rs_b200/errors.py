"""PvwError (src/errors.rs:11-73): the variant names are the reference's; the C ABI status codes map onto them 1:1."""


class PvwError(Exception):
    def __init__(self, variant: str, msg: str = ""):
        super().__init__(f"{variant}: {msg}")
        self.variant = variant
        self.msg = msg
