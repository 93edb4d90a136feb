"""Test-only stand-in for the slice of sacred that experiments/train_bnn.py and
exp_utils.py use (train_bnn.py:16-18,35-36,38-123,155; exp_utils.py:15,554-562).
sacred is not installed in this image.  Never on the product path.

Supported: `Experiment(name)`, `@ex.config` (the function body is executed and its local
variables become the configuration; values given as `config_updates` win over the
assignments in the body, so statements that depend on them see the updated value, as in
sacred), `ex.capture`, `ex.automain` / `ex.main`, `ex.observers`, `ex.add_config`,
`ex.captured_out_filter`, `ex.run(config_updates=...)` -> object with `.result` / `.config`,
and `observers.FileStorageObserver(basedir)` with `.dir`, `.run_entry`, `.save_json`.
"""
import ast
import functools
import inspect
import logging
import textwrap

from . import observers, utils  # noqa: F401

__version__ = "0.0-shim"


class _Preset(dict):
    """namespace for a config function: assignments to preset keys are ignored"""

    def __init__(self, preset):
        super().__init__(preset)
        self._preset = set(preset)

    def __setitem__(self, k, v):
        if k in self._preset:
            return
        super().__setitem__(k, v)


class Run:
    def __init__(self, experiment, config):
        self.experiment = experiment
        self.config = config
        self.observers = list(experiment.observers)
        self.result = None
        self.info = {}


class Experiment:
    def __init__(self, name="experiment", **kwargs):
        self.path = name
        self.observers = []
        self.captured_out_filter = None
        self._config_fns = []
        self._config_dicts = []
        self._main = None
        self.current_run = None
        self.logger = logging.getLogger(name)

    # -- configuration
    def config(self, fn):
        self._config_fns.append(fn)
        return fn

    def add_config(self, cfg=None, **kw):
        self._config_dicts.append(dict(cfg or {}, **kw))

    def _evaluate_config(self, updates):
        cfg = {}
        for d in self._config_dicts:
            cfg.update(d)
        cfg.update(updates)
        for fn in self._config_fns:
            src = textwrap.dedent(inspect.getsource(fn))
            tree = ast.parse(src)
            fdef = next(n for n in tree.body if isinstance(n, ast.FunctionDef))
            body = ast.Module(body=fdef.body, type_ignores=[])
            ast.increment_lineno(body, fn.__code__.co_firstlineno - 1)
            ns = _Preset({k: v for k, v in cfg.items()})
            exec(compile(body, inspect.getsourcefile(fn) or "<config>", "exec"), fn.__globals__, ns)
            for k, v in ns.items():
                if not k.startswith("_") and not inspect.ismodule(v) and not callable(v):
                    cfg[k] = v
        return cfg

    # -- captured functions
    def capture(self, fn=None, prefix=None):
        if fn is None:
            return functools.partial(self.capture, prefix=prefix)
        sig = inspect.signature(fn)

        @functools.wraps(fn)
        def wrapper(*args, **kwargs):
            run = self.current_run
            if run is None:
                raise RuntimeError("captured function called outside of a run")
            bound = sig.bind_partial(*args, **kwargs)
            for name, par in sig.parameters.items():
                if name in bound.arguments or par.kind in (par.VAR_POSITIONAL, par.VAR_KEYWORD):
                    continue
                if name == "_run":
                    bound.arguments[name] = run
                elif name == "_log":
                    bound.arguments[name] = self.logger
                elif name == "_config":
                    bound.arguments[name] = run.config
                elif name in run.config:
                    bound.arguments[name] = run.config[name]
            return fn(*bound.args, **bound.kwargs)
        return wrapper

    def main(self, fn):
        self._main = self.capture(fn)
        return self._main

    def automain(self, fn):
        captured = self.main(fn)
        if fn.__module__ == "__main__":
            self.run_commandline()
        return captured

    def run_commandline(self, argv=None):
        import sys
        argv = list(sys.argv if argv is None else argv)[1:]
        updates = {}
        if argv and argv[0] == "with":
            for item in argv[1:]:
                k, v = item.split("=", 1)
                try:
                    v = ast.literal_eval(v)
                except (ValueError, SyntaxError):
                    pass
                updates[k] = v
        return self.run(config_updates=updates)

    def run(self, command_name=None, config_updates=None, **kwargs):
        cfg = self._evaluate_config(dict(config_updates or {}))
        run = Run(self, cfg)
        run.observers = list(self.observers)     # config functions may have appended observers
        self.current_run = run
        try:
            for obs in run.observers:
                obs.started_event(ex_info={"name": self.path}, command=command_name or "main", host_info={},
                                  start_time=None, config=cfg, meta_info={}, _id=None)
            run.result = self._main()
            for obs in run.observers:
                obs.completed_event(stop_time=None, result=run.result)
        finally:
            self.current_run = None
        return run
