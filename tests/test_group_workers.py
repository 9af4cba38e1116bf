"""The enqueueing-thread pool of the device group (hexed_b200/csrc/group_workers.hpp, used by group.cu for several GPUs behind one
Kernel_mesh) is plain C++: its dispatch logic is exercised here on the CPU, since the GPU-less container cannot run the group itself
with threads (the host-thread emulation of the kernels keeps the serial loop)."""
import os
import subprocess
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PROGRAM = textwrap.dedent(r'''
    #include "group_workers.hpp"
    #include <atomic>
    #include <cstdio>
    #include <set>
    int main()
    {
      for (int n : {1, 2, 3, 8}) {
        hb::Workers w;
        w.start(n);
        if (n > 1 && !w.threaded) { std::puts("not threaded"); return 1; }
        std::vector<std::atomic<int>> hits(n);
        std::vector<std::thread::id> who(n);
        for (int rep = 0; rep < 2000; ++rep) {
          const int rc = w.run([&](int r) -> int { ++hits[r]; who[r] = std::this_thread::get_id(); return 0; });
          if (rc) { std::puts("unexpected error code"); return 2; }
        }
        for (int r = 0; r < n; ++r) if (hits[r] != 2000) { std::printf("rank %d ran %d times\n", r, int(hits[r])); return 3; }
        if (who[0] != std::this_thread::get_id()) { std::puts("rank 0 must run on the caller"); return 4; }
        if (int(std::set<std::thread::id>(who.begin(), who.end()).size()) != n) { std::puts("ranks must have their own threads"); return 5; }
        // the first non-zero code comes back, and every rank still ran (a failing rank must not leave the others half enqueued and unjoined)
        std::vector<std::atomic<int>> ran(n);
        const int rc = w.run([&](int r) -> int { ++ran[r]; return r == n - 1 ? 40 + r : 0; });
        if (rc != 40 + n - 1) { std::printf("error code %d\n", rc); return 6; }
        for (int r = 0; r < n; ++r) if (ran[r] != 1) { std::puts("a rank was skipped after an error"); return 7; }
        // results of the previous dispatch must not leak into the next one
        if (w.run([&](int) -> int { return 0; })) { std::puts("stale error code"); return 8; }
      }
      std::puts("ok");
      return 0;
    }
''')


def test_group_workers_dispatch(tmp_path):
    src = tmp_path/"workers_test.cpp"
    src.write_text(PROGRAM)
    exe = tmp_path/"workers_test"
    subprocess.run(["g++", "-std=c++17", "-O1", "-pthread", "-I", os.path.join(ROOT, "hexed_b200", "csrc"), "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == "ok", (out.returncode, out.stdout, out.stderr)


def test_group_workers_serial_switch(tmp_path):
    """HEXED_B200_GROUP_THREADS=0 keeps the serial loop"""
    src = tmp_path/"serial.cpp"
    src.write_text(textwrap.dedent(r'''
        #include "group_workers.hpp"
        #include <cstdio>
        int main() { hb::Workers w; w.start(4); int order = 0, bad = 0; w.run([&](int r) -> int { bad |= r != order++; return 0; }); std::puts(!w.threaded && !bad ? "ok" : "bad"); return 0; }
    '''))
    exe = tmp_path/"serial"
    subprocess.run(["g++", "-std=c++17", "-pthread", "-I", os.path.join(ROOT, "hexed_b200", "csrc"), "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60, env=dict(os.environ, HEXED_B200_GROUP_THREADS="0"))
    assert out.stdout.strip() == "ok"
