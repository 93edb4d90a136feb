"""FileStorageObserver of the sacred test shim: a numbered run directory under `basedir`
with run.json / config.json, `.dir`, `.run_entry`, `.save_json` (what exp_utils.sneaky_artifact,
exp_utils.py:554-562, touches)."""
import json
import os


class FileStorageObserver:
    def __init__(self, basedir, **kwargs):
        self.basedir = str(basedir)
        self.dir = None
        self.run_entry = None
        self.config = None

    def _make_run_dir(self):
        os.makedirs(self.basedir, exist_ok=True)
        ids = [int(d) for d in os.listdir(self.basedir) if d.isdigit()]
        i = max(ids, default=0) + 1
        while True:
            path = os.path.join(self.basedir, str(i))
            try:
                os.mkdir(path)
                return path
            except FileExistsError:
                i += 1

    def started_event(self, ex_info, command, host_info, start_time, config, meta_info, _id):
        self.dir = self._make_run_dir()
        self.config = config
        self.run_entry = {"experiment": dict(ex_info), "command": command, "artifacts": [], "resources": [],
                          "status": "RUNNING"}
        self.save_json(self.run_entry, "run.json")
        self.save_json({k: v for k, v in config.items() if _jsonable(v)}, "config.json")
        return os.path.basename(self.dir)

    def completed_event(self, stop_time, result):
        self.run_entry["status"] = "COMPLETED"
        self.run_entry["result"] = result if _jsonable(result) else repr(result)
        self.save_json(self.run_entry, "run.json")

    def save_json(self, obj, filename):
        with open(os.path.join(self.dir, filename), "w") as f:
            json.dump(obj, f, indent=2, sort_keys=True, default=repr)


def _jsonable(v):
    try:
        json.dumps(v)
        return True
    except TypeError:
        return False
