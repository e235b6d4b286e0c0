"""Reader for the reference's run-time parameter files (deal.II ParameterHandler ".prm" syntax,
as written by ParameterAcceptor: source/time_loop.template.h:27-215 and the sections
A - TimeLoop ... K of the modules).

    subsection A - TimeLoop
      set final time = 2.0          # comment
      subsection nested
        set key = value
      end
    end

Only parsing and typed access live here; the mapping of sections onto the B200 path is in
time_loop.py.  Unknown sections and keys are kept (and can be listed) so that a prm written
for the reference is accepted unmodified.
"""
import re


class ParameterError(ValueError):
    pass


def _normalize(name):
    """deal.II collapses runs of white space in section and entry names"""
    return " ".join(name.split())


class Section(dict):
    """entries (str -> str) and subsections (str -> Section) of one `subsection` block"""

    def __init__(self, name=""):
        super().__init__()
        self.name = name
        self.subsections = {}
        self.accessed = set()

    def subsection(self, name):
        name = _normalize(name)
        if name not in self.subsections:
            self.subsections[name] = Section(name)
        return self.subsections[name]

    # typed access with the defaults of the reference's add_parameter() calls
    def get_str(self, key, default):
        self.accessed.add(key)
        return self.get(_normalize(key), default)

    def get_float(self, key, default):
        v = self.get_str(key, None)
        if v is None:
            return float(default)
        try:
            return float(v)
        except ValueError:
            raise ParameterError("entry »%s« of subsection »%s«: »%s« is not a number" % (key, self.name, v))

    def get_int(self, key, default):
        v = self.get_str(key, None)
        if v is None:
            return int(default)
        try:
            return int(v)
        except ValueError:
            raise ParameterError("entry »%s« of subsection »%s«: »%s« is not an integer" % (key, self.name, v))

    def get_bool(self, key, default):
        v = self.get_str(key, None)
        if v is None:
            return bool(default)
        if v.lower() in ("true", "yes", "on", "1"):
            return True
        if v.lower() in ("false", "no", "off", "0"):
            return False
        raise ParameterError("entry »%s« of subsection »%s«: »%s« is not a boolean" % (key, self.name, v))

    def get_list(self, key, default, convert=float):
        """comma separated list (dealii::Point, Tensor, std::vector patterns)"""
        v = self.get_str(key, None)
        if v is None:
            return list(default)
        items = [x.strip() for x in v.split(",")]
        if items == [""]:
            return []
        try:
            return [convert(x) for x in items]
        except ValueError:
            raise ParameterError("entry »%s« of subsection »%s«: cannot convert »%s«" % (key, self.name, v))

    def unused(self, prefix=""):
        """entries that were set in the file but never read (a typo, or a feature outside the
        hot path)"""
        here = prefix + self.name
        out = [here + "/" + k for k in self if k not in {_normalize(a) for a in self.accessed}]
        for s in self.subsections.values():
            out.extend(s.unused(here + "/" if here else ""))
        return out


_SET = re.compile(r"^set\s+(.*?)\s*=\s*(.*)$")


def parse(text):
    """-> root Section"""
    root = Section("")
    stack = [root]
    pending = ""
    for lineno, raw in enumerate(text.splitlines(), 1):
        line = raw.split("#", 1)[0].rstrip()
        if pending:
            line = line.lstrip()  # deal.II joins a continued line after stripping its indentation
        if line.endswith("\\"):
            pending += line[:-1].rstrip() + " "
            continue
        line = (pending + line).strip()
        pending = ""
        if not line:
            continue
        lower = line.lower()
        if lower.startswith("subsection"):
            name = line[len("subsection"):].strip()
            if not name:
                raise ParameterError("line %d: subsection without a name" % lineno)
            stack.append(stack[-1].subsection(name))
        elif lower == "end":
            if len(stack) == 1:
                raise ParameterError("line %d: »end« without a matching »subsection«" % lineno)
            stack.pop()
        else:
            m = _SET.match(line)
            if not m:
                raise ParameterError("line %d: cannot parse »%s«" % (lineno, raw.strip()))
            stack[-1][_normalize(m.group(1))] = m.group(2).strip()
    if len(stack) != 1:
        raise ParameterError("missing »end« for subsection »%s«" % stack[-1].name)
    return root


def load(path):
    with open(path) as f:
        return parse(f.read())
