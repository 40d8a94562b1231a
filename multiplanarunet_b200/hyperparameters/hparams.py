"""YAMLHParams on PyYAML (ruamel.yaml is not available here), mirroring
mpunet/hyperparameters/hparams.py:60-248: a dict of the YAML groups that also keeps the raw YAML
STRING and edits it line-wise, so that comments and layout of train_hparams.yaml survive
`set_value` + `save_current` (the Auditor writes dim / real_space_span / n_classes back this way)."""
import os
import re

import yaml


class YAMLHParams(dict):
    def __init__(self, yaml_path, logger=None, no_log=False, no_version_control=True, **kwargs):
        dict.__init__(self)
        self.logger = logger or (lambda *a, **k: None)
        self.yaml_path = os.path.abspath(yaml_path)
        self.project_path = os.path.split(self.yaml_path)[0]
        if not os.path.exists(self.yaml_path):
            raise OSError("YAML path '%s' does not exist" % self.yaml_path)
        with open(self.yaml_path, "r") as f:
            self.string_rep = f.read()
        loaded = yaml.safe_load(self.string_rep) or {}
        self.update({k: v for k, v in loaded.items() if k[:4] != "__CB"})

    # -- groups of the raw string ------------------------------------------------------------------
    @property
    def groups(self):
        starts = [m.start() for m in re.finditer(r"^(?![ \n#])[^\n]*?:", self.string_rep, re.MULTILINE)]
        starts = starts or [0]
        bounds = [0] + starts[1:] + [len(self.string_rep)]
        return [self.string_rep[bounds[i]:bounds[i + 1]] for i in range(len(bounds) - 1)]

    def get_group(self, group_name):
        for g in self.groups:
            body = g.lstrip("\n ")
            # skip leading comment lines of the chunk
            lines = [ln for ln in body.split("\n") if ln and not ln.lstrip().startswith("#")]
            if lines and lines[0].split(":")[0].strip() == group_name:
                return g
        raise KeyError(group_name)

    def get_from_anywhere(self, key, default=None):
        found = []
        for name, group in self.items():
            if isinstance(group, dict) and key in group:
                found.append((name, group[key]))
        if len(found) > 1:
            self.logger("[ERROR] Found key '%s' in multiple groups (%s)" % (key, [f[0] for f in found]))
            return None
        return found[0][1] if found else default

    def set_value(self, subdir, name, value, overwrite=False):
        """Set hparams[subdir][name] (in memory and in the YAML string).  Existing non-null values are
        kept unless overwrite=True (hparams.py:224-240)."""
        exists = (subdir in self and isinstance(self[subdir], dict) and name in self[subdir]) if subdir else name in self
        cur = (self[subdir][name] if subdir else self[name]) if exists else None
        if not exists:
            raise AttributeError("Entry '%s' does not exist under subdir '%s'" % (name, subdir))
        if cur is None or overwrite:
            group = self.get_group(subdir) if subdir else self.string_rep
            lines = group.split("\n")
            for i, line in enumerate(lines):
                if line.lstrip().startswith(name + ":") or line.lstrip().startswith(name + " :"):
                    comment = ""
                    if " #" in line:
                        comment = "  #" + line.split(" #", 1)[1]
                    lines[i] = line.split(":")[0] + ": {}".format(value) + comment
                    break
            else:
                raise AttributeError("No field has the name '{}'".format(name))
            self.string_rep = self.string_rep.replace(group, "\n".join(lines))
            if subdir:
                self[subdir][name] = value
            else:
                self[name] = value
            return True
        return False

    def save_current(self, out_path=None):
        with open(out_path or self.yaml_path, "w") as f:
            f.write(self.string_rep)
