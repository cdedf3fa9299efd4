"""A tiny Redis look-alike (RESP2 over TCP) — just the commands RedisStore issues: SETNX, GET,
EXISTS, MGET, APPEND, INCRBY, SET, DEL, PING, plus KEYS and a non-standard _STATS (names of the
commands served). There is no redis-server in the image; this lets the RESP client be tested
against real sockets and real framing. Runs as its own process (`python fake_redis.py` prints the
port): the store's blocking calls hold the GIL, so an in-process server could never answer."""
import socket
import threading
import time


class FakeRedis:
    def __init__(self):
        self.data = {}
        self.lock = threading.Lock()
        self.sock = socket.socket()
        self.sock.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        self.sock.bind(("127.0.0.1", 0))
        self.sock.listen(64)
        self.port = self.sock.getsockname()[1]
        self.commands = []
        self._stop = False
        self.thread = threading.Thread(target=self._accept, daemon=True)
        self.thread.start()

    def close(self):
        self._stop = True
        try:
            socket.create_connection(("127.0.0.1", self.port), timeout=1).close()
        except OSError:
            pass
        self.sock.close()

    def _accept(self):
        while not self._stop:
            try:
                c, _ = self.sock.accept()
            except OSError:
                return
            threading.Thread(target=self._serve, args=(c,), daemon=True).start()

    @staticmethod
    def _readline(f):
        line = f.readline()
        if not line:
            raise EOFError
        return line[:-2]

    def _serve(self, c):
        f = c.makefile("rb")
        try:
            while True:
                head = self._readline(f)
                assert head[:1] == b"*", head
                args = []
                for _ in range(int(head[1:])):
                    n = int(self._readline(f)[1:])
                    args.append(f.read(n))
                    f.read(2)
                c.sendall(self._run(args))
        except (EOFError, OSError, ValueError):
            pass
        finally:
            c.close()

    @staticmethod
    def _bulk(v):
        return b"$-1\r\n" if v is None else b"$%d\r\n%s\r\n" % (len(v), v)

    def _run(self, args):
        cmd = args[0].upper().decode()
        with self.lock:
            self.commands.append(cmd)
            d = self.data
            if cmd == "PING":
                return b"+PONG\r\n"
            if cmd == "SETNX":
                if args[1] in d:
                    return b":0\r\n"
                d[args[1]] = args[2]
                return b":1\r\n"
            if cmd == "SET":
                d[args[1]] = args[2]
                return b"+OK\r\n"
            if cmd == "GET":
                return self._bulk(d.get(args[1]))
            if cmd == "EXISTS":
                return b":%d\r\n" % sum(1 for k in args[1:] if k in d)
            if cmd == "MGET":
                return b"*%d\r\n" % (len(args) - 1) + b"".join(self._bulk(d.get(k)) for k in args[1:])
            if cmd == "APPEND":
                d[args[1]] = d.get(args[1], b"") + args[2]
                return b":%d\r\n" % len(d[args[1]])
            if cmd == "INCRBY":
                v = int(d.get(args[1], b"0")) + int(args[2])
                d[args[1]] = str(v).encode()
                return b":%d\r\n" % v
            if cmd == "KEYS":
                return b"*%d\r\n" % len(d) + b"".join(self._bulk(k) for k in d)
            if cmd == "_STATS":
                names = sorted(set(self.commands))
                return b"*%d\r\n" % len(names) + b"".join(self._bulk(n.encode()) for n in names)
            if cmd == "DEL":
                return b":%d\r\n" % sum(1 for k in args[1:] if d.pop(k, None) is not None)
            return b"-ERR unknown command '%s'\r\n" % cmd.encode()


def resp_call(port, *args):
    """Minimal client for the test's own inspection queries: returns a list of bytes for
    array replies, bytes / int otherwise."""
    with socket.create_connection(("127.0.0.1", port), timeout=5) as c:
        c.sendall(b"*%d\r\n" % len(args) + b"".join(b"$%d\r\n%s\r\n" % (len(a), a) for a in args))
        f = c.makefile("rb")

        def one():
            line = f.readline()[:-2]
            t, rest = line[:1], line[1:]
            if t == b":":
                return int(rest)
            if t == b"$":
                n = int(rest)
                if n < 0:
                    return None
                v = f.read(n)
                f.read(2)
                return v
            if t == b"*":
                return [one() for _ in range(int(rest))]
            return rest

        return one()


if __name__ == "__main__":
    srv = FakeRedis()
    print(srv.port, flush=True)
    while True:
        time.sleep(3600)
