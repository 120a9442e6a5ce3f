"""The two rate limiters of the per-block loop (CPU only): RateLimiter.h:13-66 (per-thread budget
per second) and RateLimiterRWMixThreads.h:22-197 (byte ratio between reader and writer threads)."""
import threading
import time


def test_rate_limiter_budget_per_second(native):
    limiter = native.elb_rate_limiter_create(4 << 20)  # 4 MiB per second
    try:
        t0 = time.time()
        waits = [native.elb_rate_limiter_wait(limiter, 1 << 20) for _ in range(9)]
        elapsed = time.time() - t0
        # blocks 1-4 pass, block 5 waits for the second to end, 6-8 pass, block 9 waits again
        assert waits == [0, 0, 0, 0, 1, 0, 0, 0, 1]
        assert 1.9 <= elapsed <= 5.0  # (generous upper bound for loaded CI hosts)
        # a second without traffic resets the budget without waiting
        time.sleep(1.1)
        assert native.elb_rate_limiter_wait(limiter, 1 << 20) == 0
    finally:
        native.elb_rate_limiter_destroy(limiter)


def test_rwmix_balancer_keeps_the_read_share(native):
    pct, block = 25, 64 << 10
    balancer = native.elb_rwmix_balancer_create(pct, 2, 2, block)
    counts = {"read": 0, "write": 0}
    lock = threading.Lock()
    stop = time.time() + 1.5

    def run(kind):
        wait = native.elb_rwmix_balancer_wait_read if kind == "read" else \
            native.elb_rwmix_balancer_wait_write
        done = 0
        while time.time() < stop:
            if wait(balancer, block) < 0:
                break
            done += 1
            if kind == "write":
                time.sleep(0.0002)  # writers are the slow side, unthrottled readers would run away
        with lock:
            counts[kind] += done

    threads = [threading.Thread(target=run, args=(kind,)) for kind in ("read", "read", "write",
                                                                        "write")]
    for thread in threads:
        thread.start()
    for thread in threads[2:]:
        thread.join()  # writers end on their own
    native.elb_rwmix_balancer_interrupt(balancer)  # readers may wait for writer progress forever
    for thread in threads[:2]:
        thread.join(timeout=10)
        assert not thread.is_alive()
    native.elb_rwmix_balancer_destroy(balancer)
    share = 100.0 * counts["read"] / (counts["read"] + counts["write"])
    assert counts["write"] > 200
    assert pct - 6 <= share <= pct + 6, (share, counts)


def test_rwmix_balancer_interrupt_releases_waiters(native):
    balancer = native.elb_rwmix_balancer_create(10, 1, 1, 4096)
    results = []

    def reader():
        # without any writer progress the reader runs into its cap and has to wait
        for _ in range(100000):
            res = native.elb_rwmix_balancer_wait_read(balancer, 4096)
            if res < 0:
                results.append(res)
                return
        results.append(0)

    thread = threading.Thread(target=reader)
    thread.start()
    time.sleep(0.3)
    assert thread.is_alive()  # blocked in the balancer
    native.elb_rwmix_balancer_interrupt(balancer)
    thread.join(timeout=5)
    assert not thread.is_alive() and results == [-1]
    native.elb_rwmix_balancer_destroy(balancer)


def test_write_gate_is_exclusive_fifo_and_loses_no_wakeup(native):
    """FileWriteGate (elb_cfg::serializeBufferedWrites): futex hand-over under contention"""
    assert native.elb_write_gate_selftest(1, 1000, 0) == 0
    assert native.elb_write_gate_selftest(2, 3000, 0) == 0
    assert native.elb_write_gate_selftest(12, 500, 20) == 0     # sleepers get woken
    assert native.elb_write_gate_selftest(40, 100, 0) == 0      # more threads than cores


def test_write_gate_serves_tickets_in_order(native):
    import threading
    gate = native.elb_write_gate_create()
    order = []
    tickets = [native.elb_write_gate_take_ticket(gate) for _ in range(6)]
    assert tickets == list(range(6))

    def turn(ticket):
        native.elb_write_gate_wait_until_near(gate, ticket)
        native.elb_write_gate_wait_turn(gate, ticket)
        order.append(ticket)
        native.elb_write_gate_leave(gate)

    threads = [threading.Thread(target=turn, args=(t,)) for t in reversed(tickets)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=20)
    assert order == tickets
    native.elb_write_gate_destroy(gate)
