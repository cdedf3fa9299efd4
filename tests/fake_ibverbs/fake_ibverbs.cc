// A software stand-in for libibverbs.so.1, enough of it to run the ibverbs transport of this
// library in CI: reliable-connected queue pairs between QPs of ONE process (threads as
// ranks), memory regions with rkey / bounds checks, SEND / RECV, RDMA WRITE (with immediate)
// and RDMA READ executed synchronously inside ibv_post_send, completion queues with a
// completion channel (pipe). The image ships neither rdma-core nor an HCA; this plays the role
// tests/fake_redis.py and tests/fake_mpi/ play for their components. Build:
//   g++ -shared -fPIC -std=c++17 -I csrc tests/fake_ibverbs/fake_ibverbs.cc -o libfakeibverbs.so
// and point GLB_IBVERBS_LIB at it.
#include <arpa/inet.h>
#include <unistd.h>

#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <vector>

#include "glb/transport/ibverbs/verbs_abi.h"

namespace {

struct FakeCq;
struct FakeQp;

struct FakeMr {
  ibv_mr pub;
  char* addr;
  size_t len;
  int access;
};

struct FakeCq {
  ibv_cq pub;
  std::mutex mu;
  std::deque<ibv_wc> wcs;
  bool armed = false;
  void push(const ibv_wc& wc) {
    bool ring = false;
    {
      std::lock_guard<std::mutex> g(mu);
      wcs.push_back(wc);
      if (armed) {
        armed = false;
        ring = true;
      }
    }
    if (ring && pub.channel != nullptr) {
      char c = 1;
      (void)!::write(reinterpret_cast<int*>(pub.channel + 1)[0], &c, 1);
    }
  }
};

struct PendingSend {  // a SEND / WRITE_WITH_IMM that found no receive posted (RNR): retried on post_recv
  std::vector<char> payload;
  bool isImm;
  uint32_t imm;
  uint32_t srcQpn;
};

struct FakeQp {
  ibv_qp pub;
  std::mutex mu;
  std::deque<ibv_recv_wr> rq;
  std::deque<std::vector<ibv_sge>> rqSge;
  uint32_t destQpn = 0;
  std::deque<PendingSend> rnr;
};

struct World {
  std::mutex mu;
  std::map<uint32_t, FakeQp*> qps;
  std::map<uint32_t, FakeMr*> mrs;
  uint32_t nextQpn = 100, nextKey = 1000;
};
World& world() {
  static World w;
  return w;
}

int fakeDeviceTag = 0;
ibv_device* theDevice() { return reinterpret_cast<ibv_device*>(&fakeDeviceTag); }

void complete(FakeCq* cq, uint64_t wrId, ibv_wc_opcode op, uint32_t bytes, uint32_t qpn, ibv_wc_status st = IBV_WC_SUCCESS,
              uint32_t imm = 0, bool withImm = false) {
  ibv_wc wc;
  std::memset(&wc, 0, sizeof(wc));
  wc.wr_id = wrId;
  wc.status = st;
  wc.opcode = op;
  wc.byte_len = bytes;
  wc.qp_num = qpn;
  wc.imm_data = imm;
  wc.wc_flags = withImm ? IBV_WC_WITH_IMM : 0;
  cq->push(wc);
}

FakeMr* findMr(uint32_t rkey, uint64_t addr, size_t len, int needAccess) {
  auto& w = world();
  std::lock_guard<std::mutex> g(w.mu);
  auto it = w.mrs.find(rkey);
  if (it == w.mrs.end()) return nullptr;
  FakeMr* m = it->second;
  const uint64_t base = reinterpret_cast<uint64_t>(m->addr);
  if (addr < base || len > m->len || addr - base > m->len - len) return nullptr;
  if ((m->access & needAccess) != needAccess) return nullptr;
  return m;
}

// Deliver a message into the first posted receive of `dst`; false when none is posted.
bool deliver(FakeQp* dst, const std::vector<char>& payload, bool isImm, uint32_t imm, uint32_t srcQpn) {
  ibv_recv_wr wr;
  std::vector<ibv_sge> sge;
  {
    std::lock_guard<std::mutex> g(dst->mu);
    if (dst->rq.empty()) return false;
    wr = dst->rq.front();
    sge = dst->rqSge.front();
    dst->rq.pop_front();
    dst->rqSge.pop_front();
  }
  size_t off = 0;
  if (!isImm) {
    for (const auto& s : sge) {
      const size_t n = std::min<size_t>(s.length, payload.size() - off);
      std::memcpy(reinterpret_cast<void*>(s.addr), payload.data() + off, n);
      off += n;
      if (off >= payload.size()) break;
    }
  }
  auto* cq = reinterpret_cast<FakeCq*>(dst->pub.recv_cq);
  complete(cq, wr.wr_id, isImm ? IBV_WC_RECV_RDMA_WITH_IMM : IBV_WC_RECV, static_cast<uint32_t>(isImm ? 0 : payload.size()),
           dst->pub.qp_num, IBV_WC_SUCCESS, imm, isImm);
  (void)srcQpn;
  return true;
}

int fakePollCq(ibv_cq* cq_, int n, ibv_wc* out) {
  auto* cq = reinterpret_cast<FakeCq*>(cq_);
  std::lock_guard<std::mutex> g(cq->mu);
  int k = 0;
  while (k < n && !cq->wcs.empty()) {
    out[k++] = cq->wcs.front();
    cq->wcs.pop_front();
  }
  return k;
}

int fakeReqNotify(ibv_cq* cq_, int) {
  auto* cq = reinterpret_cast<FakeCq*>(cq_);
  bool ring = false;
  {
    std::lock_guard<std::mutex> g(cq->mu);
    if (!cq->wcs.empty()) {
      ring = true;  // completions already waiting: fire at once (real HCAs would have, too)
    } else {
      cq->armed = true;
    }
  }
  if (ring && cq->pub.channel != nullptr) {
    char c = 1;
    (void)!::write(reinterpret_cast<int*>(cq->pub.channel + 1)[0], &c, 1);
  }
  return 0;
}

int fakePostRecv(ibv_qp* qp_, ibv_recv_wr* wr, ibv_recv_wr** bad) {
  auto* qp = reinterpret_cast<FakeQp*>(qp_);
  for (; wr != nullptr; wr = wr->next) {
    PendingSend parked;
    bool haveParked = false;
    {
      std::lock_guard<std::mutex> g(qp->mu);
      qp->rq.push_back(*wr);
      qp->rqSge.emplace_back(wr->sg_list, wr->sg_list + wr->num_sge);
      if (!qp->rnr.empty()) {
        parked = std::move(qp->rnr.front());
        qp->rnr.pop_front();
        haveParked = true;
      }
    }
    if (haveParked) deliver(qp, parked.payload, parked.isImm, parked.imm, parked.srcQpn);
  }
  (void)bad;
  return 0;
}

int fakePostSend(ibv_qp* qp_, ibv_send_wr* wr, ibv_send_wr** bad) {
  auto* qp = reinterpret_cast<FakeQp*>(qp_);
  auto* scq = reinterpret_cast<FakeCq*>(qp->pub.send_cq);
  for (; wr != nullptr; wr = wr->next) {
    FakeQp* dst = nullptr;
    {
      auto& w = world();
      std::lock_guard<std::mutex> g(w.mu);
      auto it = w.qps.find(qp->destQpn);
      if (it != w.qps.end()) dst = it->second;
    }
    size_t total = 0;
    for (int i = 0; i < wr->num_sge; i++) total += wr->sg_list[i].length;
    auto gather = [&] {
      std::vector<char> p(total);
      size_t off = 0;
      for (int i = 0; i < wr->num_sge; i++) {
        std::memcpy(p.data() + off, reinterpret_cast<const void*>(wr->sg_list[i].addr), wr->sg_list[i].length);
        off += wr->sg_list[i].length;
      }
      return p;
    };
    ibv_wc_status st = IBV_WC_SUCCESS;
    ibv_wc_opcode op = IBV_WC_SEND;
    if (dst == nullptr) {
      st = static_cast<ibv_wc_status>(12);  // retry exceeded
    } else if (wr->opcode == IBV_WR_SEND) {
      auto payload = gather();
      if (!deliver(dst, payload, false, 0, qp->pub.qp_num)) {
        std::lock_guard<std::mutex> g(dst->mu);
        dst->rnr.push_back(PendingSend{std::move(payload), false, 0, qp->pub.qp_num});
      }
    } else if (wr->opcode == IBV_WR_RDMA_WRITE || wr->opcode == IBV_WR_RDMA_WRITE_WITH_IMM) {
      op = IBV_WC_RDMA_WRITE;
      FakeMr* m = findMr(wr->wr.rdma.rkey, wr->wr.rdma.remote_addr, total, IBV_ACCESS_REMOTE_WRITE);
      if (m == nullptr) {
        st = static_cast<ibv_wc_status>(10);  // remote access error
      } else {
        auto payload = gather();
        std::memcpy(reinterpret_cast<void*>(wr->wr.rdma.remote_addr), payload.data(), total);
        if (wr->opcode == IBV_WR_RDMA_WRITE_WITH_IMM && !deliver(dst, {}, true, wr->imm_data, qp->pub.qp_num)) {
          std::lock_guard<std::mutex> g(dst->mu);
          dst->rnr.push_back(PendingSend{{}, true, wr->imm_data, qp->pub.qp_num});
        }
      }
    } else if (wr->opcode == IBV_WR_RDMA_READ) {
      op = IBV_WC_RDMA_READ;
      FakeMr* m = findMr(wr->wr.rdma.rkey, wr->wr.rdma.remote_addr, total, IBV_ACCESS_REMOTE_READ);
      if (m == nullptr) {
        st = static_cast<ibv_wc_status>(10);
      } else {
        size_t off = 0;
        for (int i = 0; i < wr->num_sge; i++) {
          std::memcpy(reinterpret_cast<void*>(wr->sg_list[i].addr),
                      reinterpret_cast<const char*>(wr->wr.rdma.remote_addr) + off, wr->sg_list[i].length);
          off += wr->sg_list[i].length;
        }
      }
    } else {
      st = static_cast<ibv_wc_status>(9);
    }
    if ((wr->send_flags & IBV_SEND_SIGNALED) || st != IBV_WC_SUCCESS) {
      complete(scq, wr->wr_id, op, static_cast<uint32_t>(total), qp->pub.qp_num, st);
    }
  }
  (void)bad;
  return 0;
}

struct FakeChannel {
  ibv_comp_channel pub;
  int writeFd;  // read end is pub.fd (must directly follow `pub`: see FakeCq::push)
};

}  // namespace

extern "C" {

ibv_device** ibv_get_device_list(int* n) {
  auto** list = static_cast<ibv_device**>(std::calloc(2, sizeof(ibv_device*)));
  list[0] = theDevice();
  if (n != nullptr) *n = 1;
  return list;
}
void ibv_free_device_list(ibv_device** l) { std::free(l); }
const char* ibv_get_device_name(ibv_device*) { return "fake0"; }

ibv_context* ibv_open_device(ibv_device* d) {
  auto* c = new ibv_context();
  std::memset(c, 0, sizeof(*c));
  c->device = d;
  c->ops.poll_cq = fakePollCq;
  c->ops.req_notify_cq = fakeReqNotify;
  c->ops.post_send = fakePostSend;
  c->ops.post_recv = fakePostRecv;
  return c;
}
int ibv_close_device(ibv_context* c) {
  delete c;
  return 0;
}
ibv_pd* ibv_alloc_pd(ibv_context* c) { return reinterpret_cast<ibv_pd*>(new void*(c)); }
int ibv_dealloc_pd(ibv_pd* pd) {
  delete reinterpret_cast<void**>(pd);
  return 0;
}
ibv_mr* ibv_reg_mr(ibv_pd* pd, void* addr, size_t len, int access) {
  auto* m = new FakeMr();
  std::memset(&m->pub, 0, sizeof(m->pub));
  m->addr = static_cast<char*>(addr);
  m->len = len;
  m->access = access;
  auto& w = world();
  std::lock_guard<std::mutex> g(w.mu);
  m->pub.pd = pd;
  m->pub.addr = addr;
  m->pub.length = len;
  m->pub.lkey = m->pub.rkey = w.nextKey++;
  w.mrs[m->pub.rkey] = m;
  return &m->pub;
}
int ibv_dereg_mr(ibv_mr* mr) {
  auto* m = reinterpret_cast<FakeMr*>(mr);
  auto& w = world();
  {
    std::lock_guard<std::mutex> g(w.mu);
    w.mrs.erase(m->pub.rkey);
  }
  delete m;
  return 0;
}
ibv_comp_channel* ibv_create_comp_channel(ibv_context* c) {
  auto* ch = new FakeChannel();
  int fds[2];
  if (::pipe(fds) != 0) return nullptr;
  ch->pub.context = c;
  ch->pub.fd = fds[0];
  ch->pub.refcnt = 0;
  ch->writeFd = fds[1];
  return &ch->pub;
}
int ibv_destroy_comp_channel(ibv_comp_channel* ch_) {
  auto* ch = reinterpret_cast<FakeChannel*>(ch_);
  ::close(ch->pub.fd);
  ::close(ch->writeFd);
  delete ch;
  return 0;
}
ibv_cq* ibv_create_cq(ibv_context* c, int cqe, void* cqctx, ibv_comp_channel* ch, int) {
  auto* cq = new FakeCq();
  std::memset(&cq->pub, 0, sizeof(cq->pub));
  cq->pub.context = c;
  cq->pub.channel = ch;
  cq->pub.cq_context = cqctx;
  cq->pub.cqe = cqe;
  return &cq->pub;
}
int ibv_destroy_cq(ibv_cq* cq) {
  delete reinterpret_cast<FakeCq*>(cq);
  return 0;
}
int ibv_get_cq_event(ibv_comp_channel* ch, ibv_cq** cq, void** ctx) {
  char c;
  if (::read(ch->fd, &c, 1) != 1) return -1;
  if (cq != nullptr) *cq = nullptr;
  if (ctx != nullptr) *ctx = nullptr;
  return 0;
}
void ibv_ack_cq_events(ibv_cq*, unsigned int) {}
ibv_qp* ibv_create_qp(ibv_pd* pd, ibv_qp_init_attr* ia) {
  auto* qp = new FakeQp();
  std::memset(&qp->pub, 0, sizeof(qp->pub));
  qp->pub.pd = pd;
  qp->pub.send_cq = ia->send_cq;
  qp->pub.recv_cq = ia->recv_cq;
  qp->pub.qp_type = ia->qp_type;
  qp->pub.context = ia->send_cq->context;
  auto& w = world();
  std::lock_guard<std::mutex> g(w.mu);
  qp->pub.qp_num = w.nextQpn++;
  w.qps[qp->pub.qp_num] = qp;
  return &qp->pub;
}
int ibv_destroy_qp(ibv_qp* qp_) {
  auto* qp = reinterpret_cast<FakeQp*>(qp_);
  auto& w = world();
  {
    std::lock_guard<std::mutex> g(w.mu);
    w.qps.erase(qp->pub.qp_num);
  }
  delete qp;
  return 0;
}
int ibv_modify_qp(ibv_qp* qp_, ibv_qp_attr* a, int mask) {
  auto* qp = reinterpret_cast<FakeQp*>(qp_);
  if (mask & IBV_QP_STATE) qp->pub.state = a->qp_state;
  if (mask & IBV_QP_DEST_QPN) qp->destQpn = a->dest_qp_num;
  return 0;
}
int ibv_query_port(ibv_context*, uint8_t, ibv_port_attr* pa) {
  pa->state = IBV_PORT_ACTIVE;
  pa->max_mtu = pa->active_mtu = IBV_MTU_4096;
  pa->lid = 7;
  return 0;
}
int ibv_query_gid(ibv_context*, uint8_t, int, ibv_gid* gid) {
  std::memset(gid, 0, sizeof(*gid));
  gid->raw[15] = 1;
  return 0;
}

}  // extern "C"
