"""The master side of distributed mode against FAKE services (CPU only): a small Python HTTP
server answers the reference's endpoints (source/Common.h:200-269, HTTPServiceSWS.cpp) with
boost-ptree-style JSON, records what the master sends and lets the test check the protocol:
per-service rank offsets and data set threads (ProgArgs.cpp:3845-3848), --gpuperservice, the tree
file upload, --svcwait retries, result aggregation incl. stonewall, error propagation."""
import http.server
import json
import os
import socket
import subprocess
import threading
import time
import urllib.parse

from elbencho_b200.build import CLI_PATH

MiB = 1 << 20


def free_port():
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        return sock.getsockname()[1]


def ptree_list(name, values):
    """boost::property_tree writes lists as repeated "item" keys"""
    return '"%s": {%s}' % (name, ", ".join('"item": "%d"' % v for v in values))


class FakeService:
    def __init__(self, bytes_per_phase, elapsed_usec, fail_prepare=None, ready_after=0.0):
        self.port = free_port()
        self.bytes_per_phase = bytes_per_phase
        self.elapsed_usec = elapsed_usec
        self.fail_prepare = fail_prepare
        self.ready_at = time.time() + ready_after
        self.requests = []       # (method, path, query dict, body)
        self.prepare_trees = []
        self.uploaded = {}
        self.bench_id = ""
        self.phase_code = 0
        self.status_polls = 0
        self.server = None
        self.thread = None

    def start(self):
        svc = self

        class Handler(http.server.BaseHTTPRequestHandler):
            protocol_version = "HTTP/1.1"

            def log_message(self, *args):
                pass

            def _reply(self, code, body):
                data = body.encode()
                self.send_response(code)
                self.send_header("Content-Length", str(len(data)))
                self.end_headers()
                self.wfile.write(data)

            def _handle(self, method):
                parsed = urllib.parse.urlparse(self.path)
                query = dict(urllib.parse.parse_qsl(parsed.query, keep_blank_values=True))
                length = int(self.headers.get("Content-Length", "0"))
                body = self.rfile.read(length).decode() if length else ""
                svc.requests.append((method, parsed.path, query, body))
                if time.time() < svc.ready_at:
                    return self._reply(503, "not ready yet")
                code, text = svc.dispatch(method, parsed.path, query, body)
                self._reply(code, text)

            def do_GET(self):
                self._handle("GET")

            def do_POST(self):
                self._handle("POST")

        self.server = http.server.ThreadingHTTPServer(("127.0.0.1", self.port), Handler)
        self.thread = threading.Thread(target=self.server.serve_forever, daemon=True)
        self.thread.start()
        return self

    def stop(self):
        self.server.shutdown()
        self.server.server_close()

    def dispatch(self, method, path, query, body):
        if path == "/protocolversion":
            return 200, "3.1.1"
        if path == "/preparefile":
            self.uploaded[query.get("FileName", "")] = body
            return 200, ""
        if path == "/preparephase":
            if self.fail_prepare:
                return 400, self.fail_prepare
            tree = json.loads(body)
            self.prepare_trees.append(tree)
            return 200, json.dumps({"BenchPathType": "1", "NumBenchPaths": "1",
                                    "size": tree["size"], "block": tree["block"],
                                    "randamount": tree.get("randamount", "0"),
                                    "ErrorHistory": ""})
        if path == "/startphase":
            self.bench_id = query["BenchID"]
            self.phase_code = int(query["PhaseCode"])
            self.status_polls = 0
            return 200, ""
        if path == "/status":
            self.status_polls += 1
            done = self.status_polls >= 2
            nthreads = int(self.prepare_trees[-1]["threads"]) if self.prepare_trees else 0
            return 200, json.dumps({
                "BenchID": self.bench_id, "PhaseName": "x", "PhaseCode": str(self.phase_code),
                "NumWorkersDone": str(nthreads if done else 0), "NumWorkersDoneWithError": "0",
                "TriggerStoneWall": "true" if done else "false",
                "NumEntriesDone": "0",
                "NumBytesDone": str(self.bytes_per_phase if done else self.bytes_per_phase // 2),
                "NumIOPSDone": str((self.bytes_per_phase if done else self.bytes_per_phase // 2)
                                   // MiB),
                "CPUUtil": "7", "ElapsedSecs": "1", "NumIOLatUSec": "0", "SumIOLatUSec": "0",
                "NumEntLatUSec": "0", "SumEntLatUSec": "0", "ErrorHistory": ""})
        if path == "/benchresult":
            iops = self.bytes_per_phase // MiB
            buckets = [0] * 112
            buckets[40] = iops
            parts = ['"BenchID": "%s"' % self.bench_id, '"PhaseName": "x"',
                     '"PhaseCode": "%d"' % self.phase_code, '"NumWorkersDone": "2"',
                     '"NumWorkersDoneWithError": "0"', '"NumEntriesDone": "0"',
                     '"NumBytesDone": "%d"' % self.bytes_per_phase, '"NumIOPSDone": "%d"' % iops,
                     '"CPUUtilStoneWall": "11"', '"CPUUtil": "13"', '"TriggerStoneWall": "true"',
                     ptree_list("ElapsedUSecList", self.elapsed_usec), '"ErrorHistory": ""']
            for prefix in ("IOPS_", "Entries_"):
                num = iops if prefix == "IOPS_" else 0
                parts += ['"%sLatNumValues": "%d"' % (prefix, num),
                          '"%sLatMicroSecTotal": "%d"' % (prefix, num * 1000),
                          '"%sLatMinMicroSec": "%d"' % (prefix, 900 if num else 2 ** 64 - 1),
                          '"%sLatMaxMicroSec": "%d"' % (prefix, 1100 if num else 0),
                          ptree_list(prefix + "LatHistoList", buckets if num else [0] * 112)]
            return 200, "{" + ", ".join(parts) + "}"
        if path == "/interruptphase":
            return 200, ""
        return 404, "Unknown resource: " + path


def run_master(*args, timeout=120):
    return subprocess.run([CLI_PATH] + list(args), capture_output=True, text=True, timeout=timeout)


def table_value(stdout, phase, result_type, column=-1):
    in_phase = False
    for line in stdout.splitlines():
        if line.startswith(phase + " "):
            in_phase = True
        elif line and not line.startswith(" ") and in_phase:
            in_phase = False
        if in_phase and result_type in line and ":" in line:
            return line.split(":", 1)[1].split()[column]
    raise AssertionError("row %r of phase %r not found in:\n%s" % (result_type, phase, stdout))


def test_master_protocol_rank_offsets_and_aggregation(tmp_path):
    services = [FakeService(64 * MiB, [900000, 1000000]).start(),
                FakeService(32 * MiB, [1500000, 2000000]).start()]
    tree = tmp_path / "tree.txt"
    tree.write_text("d d1\nf 1048576 d1/a\n")
    hosts = ",".join("127.0.0.1:%d" % s.port for s in services)
    try:
        res = run_master("-w", "-r", "-t", "2", "-b", "1M", "-s", "48M", "--verify", "1", "--gpuids",
                         "3,5", "--gpuperservice", "--hosts", hosts, "--nolive", "--lat",
                         "--limitread", "7M", "--randalgo", "fast", "--treefile", str(tree),
                         "--timelimit", "50", "--svcelapsed", str(tmp_path / "bench"))
        assert res.returncode == 0, res.stdout + res.stderr
        for idx, svc in enumerate(services):
            paths = [req[1] for req in svc.requests]
            # tree file first, then prepare, then start / status... / benchresult per phase
            assert paths.index("/preparefile") < paths.index("/preparephase")
            assert svc.uploaded == {"treefile.txt": tree.read_text()}
            prep = svc.prepare_trees[0]
            assert prep["rankoffset"] == str(idx * 2)          # ProgArgs.cpp:3845-3848
            assert prep["datasetthreads"] == "4"
            assert prep["threads"] == "2" and prep["block"] == str(MiB) and prep["size"] == str(48 * MiB)
            assert prep["gpuids"] == ("3", "5")[idx]            # --gpuperservice
            assert prep["treefile"] == "treefile.txt"
            assert prep["verify"] == "1" and prep["limitread"] == str(7 * MiB)
            assert prep["randalgo"] == "fast" and prep["b200_timelimit"] == "50"
            assert prep["write"] == "true" and prep["read"] == "true"
            starts = [req for req in svc.requests if req[1] == "/startphase"]
            assert [int(req[2]["PhaseCode"]) for req in starts] == [4, 6]  # CREATEFILES, READFILES
            assert paths.count("/benchresult") == 2
            assert paths[-1] == "/interruptphase"
        # aggregation: bytes summed, first done = fastest thread, last done = slowest
        assert int(table_value(res.stdout, "WRITE", "Total MiB")) == 96
        assert table_value(res.stdout, "WRITE", "Elapsed time", 0) == "900ms"
        assert table_value(res.stdout, "WRITE", "Elapsed time", 1) == "2.000s"
        assert int(table_value(res.stdout, "WRITE", "Throughput MiB/s")) == 48
        assert "min=900us avg=1.00ms max=1.10ms" in res.stdout
        # --svcelapsed: services ordered by their slowest thread (Statistics.cpp:2079-2117)
        assert "Svc compl. time  : [ 127.0.0.1:%d=1.000s 127.0.0.1:%d=2.000s ]" % (
            services[0].port, services[1].port) in res.stdout
    finally:
        for svc in services:
            svc.stop()


def test_master_waits_for_services_and_reports_their_errors(tmp_path):
    late = FakeService(8 * MiB, [1000], ready_after=2.0).start()
    try:
        res = run_master("-w", "-t", "1", "-b", "1M", "-s", "8M", "--gpuids", "0", "--hosts",
                         "127.0.0.1:%d" % late.port, "--svcwait", "20", "--nolive",
                         str(tmp_path / "f"))
        assert res.returncode == 0, res.stdout + res.stderr
        assert sum(1 for req in late.requests if req[1] == "/status") >= 3  # incl. readiness polls
    finally:
        late.stop()
    broken = FakeService(8 * MiB, [1000], fail_prepare="Preparation phase error: no such GPU").start()
    try:
        res = run_master("-w", "-t", "1", "-b", "1M", "-s", "8M", "--gpuids", "0", "--hosts",
                         "127.0.0.1:%d" % broken.port, "--nolive", str(tmp_path / "f"))
        assert res.returncode == 1
        assert "Preparation phase error: no such GPU" in res.stderr
        assert broken.requests[-1][1] == "/interruptphase"
    finally:
        broken.stop()


def test_master_nosvcshare_and_rotatehosts(tmp_path):
    """--nosvcshare: every service gets the full data set (rank offset 0, data set threads =
    threads; ProgArgs.cpp:1288, 3845); --rotatehosts 1: the services swap ranks between phases,
    which takes a new preparation on all of them (Coordinator.cpp:382-404)"""
    services = [FakeService(16 * MiB, [1000000]).start() for _ in range(3)]
    hosts = ",".join("127.0.0.1:%d" % s.port for s in services)
    try:
        res = run_master("-w", "-t", "2", "-b", "1M", "-s", "16M", "--gpuids", "0", "--hosts", hosts,
                         "--nolive", "--nosvcshare", "--svcupint", "50", str(tmp_path / "f"))
        assert res.returncode == 0, res.stdout + res.stderr
        for svc in services:
            assert svc.prepare_trees[0]["rankoffset"] == "0"
            assert svc.prepare_trees[0]["datasetthreads"] == "2"
            svc.requests.clear()
            svc.prepare_trees.clear()

        res = run_master("-w", "-r", "-t", "2", "-b", "1M", "-s", "16M", "--gpuids", "0", "--hosts",
                         hosts, "--nolive", "--rotatehosts", "1", "--svcupint", "50",
                         str(tmp_path / "f"))
        assert res.returncode == 0, res.stdout + res.stderr
        # two preparations per service; rank offsets rotate by one position of the hosts list
        first = [svc.prepare_trees[0]["rankoffset"] for svc in services]
        second = [svc.prepare_trees[1]["rankoffset"] for svc in services]
        assert first == ["0", "2", "4"]
        assert second == ["4", "0", "2"]
        for svc in services:
            assert len(svc.prepare_trees) == 2
            assert svc.prepare_trees[1]["datasetthreads"] == "6"
    finally:
        for svc in services:
            svc.stop()


def test_help_variants_and_accepted_noops():
    for flag in ("--help", "--help-all", "--help-dist", "--help-multi", "--help-large",
                 "--help-bdev"):
        res = run_master(flag)
        assert res.returncode == 0 and "--gpuids" in res.stdout and "--treefile" in res.stdout
    res = run_master("--dryrun", "-w", "-s", "1M", "--gpuids", "0", "--cufile", "--cufiledriveropen",
                     "--svcping", "--althttpsvc", "/tmp/elb_dry_noop")
    assert res.returncode == 0, res.stderr


def test_service_password_file_both_sides(tmp_path):
    """--svcpwfile: the master sends HashTk::simple128 of the file's first line as PwHash with
    /preparephase and /preparefile; a service rejects other hashes (HTTPServiceSWS.cpp:287-296)"""
    import http.client
    pwfile = tmp_path / "pw.txt"
    pwfile.write_text("secret\nsecond line is ignored\n")
    want_hash = "81cae49641cc8a331b2199b5f77da014"  # tests/golden: simple128("secret")

    fake = FakeService(8 * MiB, [1000]).start()
    try:
        res = run_master("-w", "-t", "1", "-b", "1M", "-s", "8M", "--gpuids", "0", "--hosts",
                         "127.0.0.1:%d" % fake.port, "--svcpwfile", str(pwfile), "--nolive",
                         str(tmp_path / "f"))
        assert res.returncode == 0, res.stdout + res.stderr
        prep = [req for req in fake.requests if req[1] == "/preparephase"][0]
        assert prep[2]["PwHash"] == want_hash
    finally:
        fake.stop()

    port = free_port()
    svc = subprocess.Popen([CLI_PATH, "--service", "--foreground", "--port", str(port),
                            "--svcpwfile", str(pwfile)], stdout=subprocess.PIPE,
                           stderr=subprocess.STDOUT, text=True)
    try:
        conn = None
        for _ in range(100):
            try:
                conn = http.client.HTTPConnection("127.0.0.1", port, timeout=5)
                conn.request("GET", "/protocolversion")
                assert conn.getresponse().read() == b"3.1.1"
                break
            except OSError:
                time.sleep(0.1)

        def post(path, body):
            conn.request("POST", path, body=body)
            response = conn.getresponse()
            return response.status, response.read()

        status, body = post("/preparefile?ProtocolVersion=3.1.1&FileName=t.txt&PwHash=", "x")
        assert status == 400 and b"Invalid authorization code." in body
        status, body = post("/preparephase?ProtocolVersion=3.1.1&PwHash=0123", "{}")
        assert status == 400 and b"Invalid authorization code." in body
        status, body = post("/preparephase?ProtocolVersion=3.1.1", "{}")
        assert status == 400 and b"Missing parameter: PwHash" in body
        status, body = post("/preparefile?ProtocolVersion=3.1.1&FileName=t.txt&PwHash=" + want_hash,
                            "d x\n")
        assert status == 200
        import getpass
        os.unlink("/var/tmp/elbencho-b200_%s_p%d/t.txt" % (getpass.getuser(), port))
        os.rmdir("/var/tmp/elbencho-b200_%s_p%d" % (getpass.getuser(), port))
        conn.close()
    finally:
        run_master("--quit", "--hosts", "127.0.0.1:%d" % port)
        try:
            svc.wait(timeout=20)
        except subprocess.TimeoutExpired:
            svc.kill()

    empty = tmp_path / "empty.txt"
    empty.write_text("\n")
    res = run_master("-w", "-s", "1M", "--gpuids", "0", "--svcpwfile", str(empty), "/tmp/x")
    assert res.returncode == 1 and "First line in service password file is empty" in res.stderr


def test_master_live_csv_with_per_service_rows(tmp_path):
    """--livecsv / --livecsvex of a distributed run: totals of all services per update, then one
    line per service with threads left, CPU utilisation and host (Statistics.cpp:3017-3230)"""
    services = [FakeService(64 * MiB, [1000000]).start(), FakeService(32 * MiB, [1000000]).start()]
    hosts = ",".join("127.0.0.1:%d" % s.port for s in services)
    live_csv = tmp_path / "live.csv"
    try:
        res = run_master("-w", "-t", "2", "-b", "1M", "-s", "48M", "--gpuids", "0", "--hosts", hosts,
                         "--nolive", "--svcupint", "50", "--livecsv", str(live_csv), "--livecsvex",
                         "--label", "my,label", str(tmp_path / "f"))
        assert res.returncode == 0, res.stdout + res.stderr
    finally:
        for svc in services:
            svc.stop()
    lines = live_csv.read_text().splitlines()
    assert lines[0] == ("ISO Date,Label,Phase,RuntimeMS,Rank,MixType,Done%,DoneBytes,MiB/s,IOPS,"
                        "Entries,Entries/s,Lat Ent us,Lat IO us,Active,CPU,Service,")
    rows = [line.split(",") for line in lines[1:]]
    assert len(rows) >= 3 and all(len(row) == 18 for row in rows)
    total, first, second = rows[0], rows[1], rows[2]
    assert total[1] == "my label" and total[2] == "WRITE" and total[4] == "Total"
    # first poll: both fake services report half of their bytes and no thread done yet
    assert int(total[7]) == 32 * MiB + 16 * MiB
    # shared paths: the 48 MiB file is the whole data set of the 4 data set threads
    assert int(total[6]) == 100
    assert total[14] == "4" and total[15] == "7"
    assert first[4] == "0" and int(first[7]) == 32 * MiB and first[14] == "2"
    assert first[16] == "127.0.0.1:%d" % services[0].port
    assert second[4] == "1" and int(second[7]) == 16 * MiB
    assert second[16] == "127.0.0.1:%d" % services[1].port
    assert first[8] == "" and first[9] == ""


def test_master_result_files_and_worker_error(tmp_path):
    """CSV / JSON / text result files of a distributed run, and a service that reports a worker
    error while the phase runs (RemoteWorker.cpp:429-572: the error history of the service is
    shown and the run fails)"""
    services = [FakeService(64 * MiB, [900000, 1000000]).start(),
                FakeService(32 * MiB, [1500000, 2000000]).start()]
    hosts = ",".join("127.0.0.1:%d" % s.port for s in services)
    csv_path, json_path, txt_path = (tmp_path / n for n in ("r.csv", "r.json", "r.txt"))
    try:
        res = run_master("-w", "-t", "2", "-b", "1M", "-s", "48M", "--gpuids", "0", "--hosts", hosts,
                         "--nolive", "--label", "dist run", "--csvfile", str(csv_path),
                         "--jsonfile", str(json_path), "--resfile", str(txt_path),
                         str(tmp_path / "f"))
        assert res.returncode == 0, res.stdout + res.stderr
    finally:
        for svc in services:
            svc.stop()
    labels, values = [line.split(",") for line in csv_path.read_text().splitlines()[:2]]
    row = dict(zip(labels, values))
    assert row["label"] == "dist run" and row["operation"] == "WRITE"
    assert row["hosts"] == "2" and row["threads"] == "2"
    assert row["MiB [last]"] == "96" and row["time ms [first]"] == "900"
    assert row["time ms [last]"] == "2000" and row["MiB/s [last]"] == "48"
    doc = json.loads(json_path.read_text())
    assert doc["phase_type"] == "WRITE" and doc["label"] == "dist run"
    assert doc["last_done"]["elapsed_time_ms"] == "2000" and doc["first_done"]["elapsed_time_ms"] == "900"
    assert doc["last_done"]["bytes"] == str(96 * MiB)
    assert "WRITE" in txt_path.read_text() and "Total MiB" in txt_path.read_text()

    class FailingService(FakeService):
        def dispatch(self, method, path, query, body):
            code, text = super().dispatch(method, path, query, body)
            if path == "/status":
                tree = json.loads(text)
                tree["NumWorkersDoneWithError"] = "1"
                tree["ErrorHistory"] = "File write failed. Path: /data/f; SysErr: No space left"
                return 200, json.dumps(tree)
            return code, text

    broken = FailingService(8 * MiB, [1000]).start()
    try:
        res = run_master("-w", "-t", "1", "-b", "1M", "-s", "8M", "--gpuids", "0", "--hosts",
                         "127.0.0.1:%d" % broken.port, "--nolive", str(tmp_path / "f"))
        assert res.returncode == 1
        assert "File write failed. Path: /data/f; SysErr: No space left" in res.stderr
        assert broken.requests[-1][1] == "/interruptphase"
    finally:
        broken.stop()


def test_master_stops_after_an_expired_time_limit(tmp_path):
    """the services apply the limit themselves (they get it as b200_timelimit) and report done;
    the master prints the phase and does not start the next one (Coordinator.cpp:234-241)"""

    class SlowService(FakeService):
        def dispatch(self, method, path, query, body):
            if path == "/startphase":
                self.started_at = time.time()
            if path == "/status" and self.prepare_trees:
                limit = int(self.prepare_trees[-1]["b200_timelimit"])
                code, text = super().dispatch(method, path, query, body)
                tree = json.loads(text)
                if time.time() - self.started_at < limit:  # busy until the limit is over
                    tree["NumWorkersDone"] = "0"
                    tree["TriggerStoneWall"] = "false"
                else:
                    tree["NumWorkersDone"] = self.prepare_trees[-1]["threads"]
                    tree["TriggerStoneWall"] = "true"
                return 200, json.dumps(tree)
            return super().dispatch(method, path, query, body)

    svc = SlowService(16 * MiB, [1000000]).start()
    try:
        t0 = time.time()
        res = run_master("-w", "-r", "-t", "2", "-b", "1M", "-s", "16M", "--gpuids", "0", "--hosts",
                         "127.0.0.1:%d" % svc.port, "--nolive", "--timelimit", "1", "--svcupint",
                         "50", str(tmp_path / "f"))
        assert res.returncode == 0, res.stdout + res.stderr
        assert time.time() - t0 >= 1.0
        assert "Terminating due to phase time limit." in res.stdout
        assert "WRITE" in res.stdout and "\nREAD " not in res.stdout
        starts = [req for req in svc.requests if req[1] == "/startphase"]
        assert len(starts) == 1
    finally:
        svc.stop()
