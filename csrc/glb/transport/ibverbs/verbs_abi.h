// Minimal mirror of the libibverbs ABI (rdma-core, ABI version 1 — unchanged since 2005 apart
// from fields appended at the end of structs) needed by this transport. The image has no
// rdma-core headers, and libibverbs is loaded with dlopen at run time, so the handful of
// structs, enums and the `ibv_context_ops` dispatch table that verbs.h's inline fast-path
// functions (ibv_post_send / ibv_post_recv / ibv_poll_cq / ibv_req_notify_cq) go through are
// declared here. Only fields up to the last one this transport reads are spelled out.
// Layout source: rdma-core libibverbs/verbs.h (struct definitions are part of its stable ABI).
#pragma once

#include <pthread.h>

#include <cstddef>
#include <cstdint>
#include <string>

extern "C" {

struct ibv_device;
struct ibv_pd;
struct ibv_srq;
struct ibv_mw;
struct ibv_mw_bind;
struct ibv_ah;
struct ibv_comp_channel {
  struct ibv_context* context;
  int fd;
  int refcnt;
};

union ibv_gid {
  uint8_t raw[16];
  struct {
    uint64_t subnet_prefix;
    uint64_t interface_id;
  } global;
};

enum ibv_mtu { IBV_MTU_256 = 1, IBV_MTU_512 = 2, IBV_MTU_1024 = 3, IBV_MTU_2048 = 4, IBV_MTU_4096 = 5 };
enum ibv_port_state { IBV_PORT_NOP = 0, IBV_PORT_DOWN = 1, IBV_PORT_INIT = 2, IBV_PORT_ARMED = 3, IBV_PORT_ACTIVE = 4 };

// Prefix shared by the current and the "compat" port attribute structs.
struct ibv_port_attr {
  enum ibv_port_state state;
  enum ibv_mtu max_mtu;
  enum ibv_mtu active_mtu;
  int gid_tbl_len;
  uint32_t port_cap_flags;
  uint32_t max_msg_sz;
  uint32_t bad_pkey_cntr;
  uint32_t qkey_viol_cntr;
  uint16_t pkey_tbl_len;
  uint16_t lid;
  uint16_t sm_lid;
  uint8_t lmc;
  uint8_t max_vl_num;
  uint8_t sm_sl;
  uint8_t subnet_timeout;
  uint8_t init_type_reply;
  uint8_t active_width;
  uint8_t active_speed;
  uint8_t phys_state;
  uint8_t link_layer;
  uint8_t flags;
  uint16_t port_cap_flags2;
  uint32_t active_speed_ex;
  uint8_t reserved_for_growth[64];  // never read: keeps newer libraries from writing past the struct
};

enum ibv_access_flags {
  IBV_ACCESS_LOCAL_WRITE = 1,
  IBV_ACCESS_REMOTE_WRITE = 1 << 1,
  IBV_ACCESS_REMOTE_READ = 1 << 2,
  IBV_ACCESS_REMOTE_ATOMIC = 1 << 3,
};

struct ibv_mr {
  struct ibv_context* context;
  struct ibv_pd* pd;
  void* addr;
  size_t length;
  uint32_t handle;
  uint32_t lkey;
  uint32_t rkey;
};

enum ibv_wc_status { IBV_WC_SUCCESS = 0 };
enum ibv_wc_opcode {
  IBV_WC_SEND = 0,
  IBV_WC_RDMA_WRITE = 1,
  IBV_WC_RDMA_READ = 2,
  IBV_WC_RECV = 1 << 7,
  IBV_WC_RECV_RDMA_WITH_IMM = (1 << 7) + 1,
};
enum ibv_wc_flags { IBV_WC_GRH = 1, IBV_WC_WITH_IMM = 2 };

struct ibv_wc {
  uint64_t wr_id;
  enum ibv_wc_status status;
  enum ibv_wc_opcode opcode;
  uint32_t vendor_err;
  uint32_t byte_len;
  uint32_t imm_data;  // network byte order
  uint32_t qp_num;
  uint32_t src_qp;
  unsigned int wc_flags;
  uint16_t pkey_index;
  uint16_t slid;
  uint8_t sl;
  uint8_t dlid_path_bits;
};

struct ibv_sge {
  uint64_t addr;
  uint32_t length;
  uint32_t lkey;
};

enum ibv_wr_opcode {
  IBV_WR_RDMA_WRITE = 0,
  IBV_WR_RDMA_WRITE_WITH_IMM = 1,
  IBV_WR_SEND = 2,
  IBV_WR_SEND_WITH_IMM = 3,
  IBV_WR_RDMA_READ = 4,
};
enum ibv_send_flags { IBV_SEND_FENCE = 1, IBV_SEND_SIGNALED = 2, IBV_SEND_SOLICITED = 4, IBV_SEND_INLINE = 8 };

struct ibv_mw_bind_info {
  struct ibv_mr* mr;
  uint64_t addr;
  uint64_t length;
  unsigned int mw_access_flags;
};

struct ibv_send_wr {
  uint64_t wr_id;
  struct ibv_send_wr* next;
  struct ibv_sge* sg_list;
  int num_sge;
  enum ibv_wr_opcode opcode;
  unsigned int send_flags;
  uint32_t imm_data;  // network byte order
  union {
    struct {
      uint64_t remote_addr;
      uint32_t rkey;
    } rdma;
    struct {
      uint64_t remote_addr;
      uint64_t compare_add;
      uint64_t swap;
      uint32_t rkey;
    } atomic;
    struct {
      struct ibv_ah* ah;
      uint32_t remote_qpn;
      uint32_t remote_qkey;
    } ud;
  } wr;
  union {
    struct {
      uint32_t remote_srqn;
    } xrc;
  } qp_type;
  union {
    struct {
      struct ibv_mw* mw;
      uint32_t rkey;
      struct ibv_mw_bind_info bind_info;
    } bind_mw;
    struct {
      void* hdr;
      uint16_t hdr_sz;
      uint16_t mss;
    } tso;
  };
};
static_assert(sizeof(struct ibv_send_wr) == 128, "ibv_send_wr layout");

struct ibv_recv_wr {
  uint64_t wr_id;
  struct ibv_recv_wr* next;
  struct ibv_sge* sg_list;
  int num_sge;
};

enum ibv_qp_type { IBV_QPT_RC = 2, IBV_QPT_UC = 3, IBV_QPT_UD = 4 };
enum ibv_qp_state { IBV_QPS_RESET = 0, IBV_QPS_INIT = 1, IBV_QPS_RTR = 2, IBV_QPS_RTS = 3, IBV_QPS_ERR = 6 };
enum ibv_mig_state { IBV_MIG_MIGRATED = 0 };

struct ibv_qp_cap {
  uint32_t max_send_wr;
  uint32_t max_recv_wr;
  uint32_t max_send_sge;
  uint32_t max_recv_sge;
  uint32_t max_inline_data;
};

struct ibv_cq {
  struct ibv_context* context;
  struct ibv_comp_channel* channel;
  void* cq_context;
  uint32_t handle;
  int cqe;
  // (mutex, cond, event counters follow; never touched here)
};

struct ibv_qp_init_attr {
  void* qp_context;
  struct ibv_cq* send_cq;
  struct ibv_cq* recv_cq;
  struct ibv_srq* srq;
  struct ibv_qp_cap cap;
  enum ibv_qp_type qp_type;
  int sq_sig_all;
};

struct ibv_global_route {
  union ibv_gid dgid;
  uint32_t flow_label;
  uint8_t sgid_index;
  uint8_t hop_limit;
  uint8_t traffic_class;
};

struct ibv_ah_attr {
  struct ibv_global_route grh;
  uint16_t dlid;
  uint8_t sl;
  uint8_t src_path_bits;
  uint8_t static_rate;
  uint8_t is_global;
  uint8_t port_num;
};

struct ibv_qp_attr {
  enum ibv_qp_state qp_state;
  enum ibv_qp_state cur_qp_state;
  enum ibv_mtu path_mtu;
  enum ibv_mig_state path_mig_state;
  uint32_t qkey;
  uint32_t rq_psn;
  uint32_t sq_psn;
  uint32_t dest_qp_num;
  unsigned int qp_access_flags;
  struct ibv_qp_cap cap;
  struct ibv_ah_attr ah_attr;
  struct ibv_ah_attr alt_ah_attr;
  uint16_t pkey_index;
  uint16_t alt_pkey_index;
  uint8_t en_sqd_async_notify;
  uint8_t sq_draining;
  uint8_t max_rd_atomic;
  uint8_t max_dest_rd_atomic;
  uint8_t min_rnr_timer;
  uint8_t port_num;
  uint8_t timeout;
  uint8_t retry_cnt;
  uint8_t rnr_retry;
  uint8_t alt_port_num;
  uint8_t alt_timeout;
  uint32_t rate_limit;
};

enum ibv_qp_attr_mask {
  IBV_QP_STATE = 1 << 0,
  IBV_QP_ACCESS_FLAGS = 1 << 3,
  IBV_QP_PKEY_INDEX = 1 << 4,
  IBV_QP_PORT = 1 << 5,
  IBV_QP_AV = 1 << 7,
  IBV_QP_PATH_MTU = 1 << 8,
  IBV_QP_TIMEOUT = 1 << 9,
  IBV_QP_RETRY_CNT = 1 << 10,
  IBV_QP_RNR_RETRY = 1 << 11,
  IBV_QP_RQ_PSN = 1 << 12,
  IBV_QP_MAX_QP_RD_ATOMIC = 1 << 13,
  IBV_QP_MIN_RNR_TIMER = 1 << 15,
  IBV_QP_SQ_PSN = 1 << 16,
  IBV_QP_MAX_DEST_RD_ATOMIC = 1 << 17,
  IBV_QP_DEST_QPN = 1 << 20,
};

struct ibv_qp {
  struct ibv_context* context;
  void* qp_context;
  struct ibv_pd* pd;
  struct ibv_cq* send_cq;
  struct ibv_cq* recv_cq;
  struct ibv_srq* srq;
  uint32_t handle;
  uint32_t qp_num;
  enum ibv_qp_state state;
  enum ibv_qp_type qp_type;
  // (mutex, cond, events_completed follow)
};

// The dispatch table inside ibv_context: the data-path verbs are function pointers here
// (verbs.h wraps them in static inline functions, which therefore are not exported symbols).
struct ibv_context_ops {
  void* _compat_query_device;
  void* _compat_query_port;
  void* _compat_alloc_pd;
  void* _compat_dealloc_pd;
  void* _compat_reg_mr;
  void* _compat_rereg_mr;
  void* _compat_dereg_mr;
  void* alloc_mw;
  void* bind_mw;
  void* dealloc_mw;
  void* _compat_create_cq;
  int (*poll_cq)(struct ibv_cq* cq, int num_entries, struct ibv_wc* wc);
  int (*req_notify_cq)(struct ibv_cq* cq, int solicited_only);
  void* _compat_cq_event;
  void* _compat_resize_cq;
  void* _compat_destroy_cq;
  void* _compat_create_srq;
  void* _compat_modify_srq;
  void* _compat_query_srq;
  void* _compat_destroy_srq;
  void* post_srq_recv;
  void* _compat_create_qp;
  void* _compat_query_qp;
  void* _compat_modify_qp;
  void* _compat_destroy_qp;
  int (*post_send)(struct ibv_qp* qp, struct ibv_send_wr* wr, struct ibv_send_wr** bad_wr);
  int (*post_recv)(struct ibv_qp* qp, struct ibv_recv_wr* wr, struct ibv_recv_wr** bad_wr);
  void* _compat_create_ah;
  void* _compat_destroy_ah;
  void* _compat_attach_mcast;
  void* _compat_detach_mcast;
  void* _compat_async_event;
};

struct ibv_context {
  struct ibv_device* device;
  struct ibv_context_ops ops;
  int cmd_fd;
  int async_fd;
  int num_comp_vectors;
  pthread_mutex_t mutex;
  void* abi_compat;
};

}  // extern "C"

namespace glb {
namespace transport {
namespace ibverbs {

// Exported entry points of libibverbs.so.1, resolved with dlsym.
struct VerbsApi {
  struct ibv_device** (*get_device_list)(int*);
  void (*free_device_list)(struct ibv_device**);
  const char* (*get_device_name)(struct ibv_device*);
  struct ibv_context* (*open_device)(struct ibv_device*);
  int (*close_device)(struct ibv_context*);
  struct ibv_pd* (*alloc_pd)(struct ibv_context*);
  int (*dealloc_pd)(struct ibv_pd*);
  struct ibv_mr* (*reg_mr)(struct ibv_pd*, void*, size_t, int);
  int (*dereg_mr)(struct ibv_mr*);
  struct ibv_comp_channel* (*create_comp_channel)(struct ibv_context*);
  int (*destroy_comp_channel)(struct ibv_comp_channel*);
  struct ibv_cq* (*create_cq)(struct ibv_context*, int, void*, struct ibv_comp_channel*, int);
  int (*destroy_cq)(struct ibv_cq*);
  int (*get_cq_event)(struct ibv_comp_channel*, struct ibv_cq**, void**);
  void (*ack_cq_events)(struct ibv_cq*, unsigned int);
  struct ibv_qp* (*create_qp)(struct ibv_pd*, struct ibv_qp_init_attr*);
  int (*destroy_qp)(struct ibv_qp*);
  int (*modify_qp)(struct ibv_qp*, struct ibv_qp_attr*, int);
  int (*query_port)(struct ibv_context*, uint8_t, struct ibv_port_attr*);
  int (*query_gid)(struct ibv_context*, uint8_t, int, union ibv_gid*);
};

// Loads the library (GLB_IBVERBS_LIB, else libibverbs.so.1) once; nullptr + reason when unavailable.
const VerbsApi* verbs(std::string* why = nullptr);

}  // namespace ibverbs
}  // namespace transport
}  // namespace glb
