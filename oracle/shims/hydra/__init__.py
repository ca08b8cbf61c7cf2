"""TEST INFRASTRUCTURE -- stand-in for `hydra` (absent): the reference only uses the
`@hydra.main(...)` decorator on its `main(cfg)` entry points, which the golden
generator never calls (it builds the cfg from the YAML itself)."""


def main(version_base=None, config_path=None, config_name=None):
    def decorator(fn):
        return fn
    return decorator
