"""Test helper: a separate interpreter that loads scene files with both loaders - the library's ygl_scene_load and the
reference's load_scene (oracle/_ref) - and reports whether they agree. The differential tests feed it files the
reference may crash on (it reads out of bounds on some inputs it accepts); a crash ends this process, not pytest, and the
caller starts a new one. Protocol: one path per line on stdin, one verdict per line on stdout:
"same" | "refused" | "accepted by one loader only (...)" | "scenes differ at ..."."""
import os
import select
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def verdict(ref, lib, assert_scenes_identical, path):
    try:
        ours, ours_error = lib.load_scene(path), None
    except lib.YglError as e:
        ours, ours_error = None, str(e)
    try:
        theirs, theirs_error = ref.load_scene(path), None
    except RuntimeError as e:
        theirs, theirs_error = None, str(e)
    if (ours is None) != (theirs is None):
        return f"accepted by one loader only (ours: {ours_error!r}, reference: {theirs_error!r})"
    if ours is None:
        return "refused"
    try:
        assert_scenes_identical(ours, theirs)
    except AssertionError as e:
        return "scenes differ at " + str(e).splitlines()[0]
    return "same"


def main():
    for p in (os.path.join(HERE, "..", "yocto-gl_b200"), os.path.join(HERE, "..", "oracle"), HERE):
        sys.path.insert(0, p)
    import refbind
    from ygl_b200 import lib
    from test_sceneio import assert_scenes_identical
    ref = refbind.Ref()
    print("ready", flush=True)
    for line in sys.stdin:
        print(verdict(ref, lib, assert_scenes_identical, line.rstrip("\n")).replace("\n", " "), flush=True)


class LoaderPair:
    """verdict(path) through a worker process that is restarted when it dies or stalls"""

    def __init__(self, timeout=60.0):
        self.timeout, self.proc = timeout, None

    def _start(self):
        self.proc = subprocess.Popen([sys.executable, os.path.abspath(__file__)], stdin=subprocess.PIPE, stdout=subprocess.PIPE,
                                     stderr=subprocess.DEVNULL, text=True, bufsize=1)
        if self._line(180.0) != "ready":
            raise RuntimeError("loader worker did not start")

    def _line(self, timeout):
        ready, _, _ = select.select([self.proc.stdout], [], [], timeout)
        if not ready:
            self.proc.kill()
            return None
        line = self.proc.stdout.readline()
        return line.rstrip("\n") if line else None

    def verdict(self, path):
        if self.proc is None or self.proc.poll() is not None:
            self._start()
        self.proc.stdin.write(str(path) + "\n")
        self.proc.stdin.flush()
        line = self._line(self.timeout)
        if line is None:                      # the worker died (or stalled and was killed) on this file
            self.proc.wait()
            self.proc = None
            return "reference crashed"
        return line

    def close(self):
        if self.proc is not None and self.proc.poll() is None:
            self.proc.stdin.close()
            self.proc.wait(timeout=30)
        self.proc = None


if __name__ == "__main__":
    main()
