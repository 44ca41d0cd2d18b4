// Rendezvous between the rank processes of one communicator.
//
// Rank 0 listens on an abstract-namespace UNIX socket named after the session;
// every other rank connects to it.  All host-side exchanges (byte blobs, POSIX
// fds of cuMem allocations via SCM_RIGHTS, barriers) go through that hub.  This
// replaces the ssh/sshd-port-23 process wiring the reference's multi-instance
// containers use (convoy/settings.py:4391-4443) with plain local IPC.
#include <errno.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/socket.h>
#include <sys/un.h>
#include <time.h>
#include <unistd.h>
#include <poll.h>
#include "internal.h"

static thread_local char g_err[512];
void sy_set_error(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap);
  if (getenv("SHIPYARD_COLL_DEBUG")) fprintf(stderr, "[shipyard-coll] %s\n", g_err);
}
extern "C" const char* sy_last_error(void) { return g_err; }

struct Hub {
  int rank = 0, world = 1;
  int listen_fd = -1;
  std::vector<int> conn;  // rank 0: conn[r] for r>0 ; others: conn[0] = socket to hub
};

static int full_write(int fd, const void* p, size_t n) {
  const char* b = (const char*)p;
  while (n) {
    ssize_t w = ::send(fd, b, n, MSG_NOSIGNAL);
    if (w < 0) { if (errno == EINTR) continue; return -1; }
    b += w; n -= (size_t)w;
  }
  return 0;
}
static int full_read(int fd, void* p, size_t n) {
  char* b = (char*)p;
  while (n) {
    ssize_t r = ::recv(fd, b, n, 0);
    if (r < 0) { if (errno == EINTR) continue; return -1; }
    if (r == 0) return -1;
    b += r; n -= (size_t)r;
  }
  return 0;
}
static int send_fd(int sock, int fd) {
  char dummy = 'F';
  struct iovec iov = {&dummy, 1};
  char ctrl[CMSG_SPACE(sizeof(int))];
  memset(ctrl, 0, sizeof ctrl);
  struct msghdr msg = {};
  msg.msg_iov = &iov; msg.msg_iovlen = 1;
  msg.msg_control = ctrl; msg.msg_controllen = sizeof ctrl;
  struct cmsghdr* cm = CMSG_FIRSTHDR(&msg);
  cm->cmsg_level = SOL_SOCKET; cm->cmsg_type = SCM_RIGHTS; cm->cmsg_len = CMSG_LEN(sizeof(int));
  memcpy(CMSG_DATA(cm), &fd, sizeof(int));
  for (;;) {
    ssize_t r = sendmsg(sock, &msg, MSG_NOSIGNAL);
    if (r < 0 && errno == EINTR) continue;
    return r == 1 ? 0 : -1;
  }
}
static int recv_fd(int sock, int* fd) {
  char dummy;
  struct iovec iov = {&dummy, 1};
  char ctrl[CMSG_SPACE(sizeof(int))];
  struct msghdr msg = {};
  msg.msg_iov = &iov; msg.msg_iovlen = 1;
  msg.msg_control = ctrl; msg.msg_controllen = sizeof ctrl;
  for (;;) {
    ssize_t r = recvmsg(sock, &msg, 0);
    if (r < 0 && errno == EINTR) continue;
    if (r != 1) return -1;
    break;
  }
  struct cmsghdr* cm = CMSG_FIRSTHDR(&msg);
  if (!cm || cm->cmsg_type != SCM_RIGHTS) return -1;
  memcpy(fd, CMSG_DATA(cm), sizeof(int));
  return 0;
}

static socklen_t make_addr(struct sockaddr_un* a, const std::string& session) {
  memset(a, 0, sizeof *a);
  a->sun_family = AF_UNIX;
  // abstract namespace: leading NUL, no filesystem residue to clean up
  std::string name = "shipyard-coll-" + session;
  if (name.size() > sizeof(a->sun_path) - 2) name.resize(sizeof(a->sun_path) - 2);
  memcpy(a->sun_path + 1, name.data(), name.size());
  return (socklen_t)(offsetof(struct sockaddr_un, sun_path) + 1 + name.size());
}

static double now_s() {
  struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + ts.tv_nsec * 1e-9;
}

Hub* hub_create(int rank, int world, const std::string& session, int timeout_ms) {
  Hub* h = new Hub();
  h->rank = rank; h->world = world;
  if (world == 1) return h;
  struct sockaddr_un addr; socklen_t alen = make_addr(&addr, session);
  double deadline = now_s() + timeout_ms * 1e-3;
  if (rank == 0) {
    h->listen_fd = socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
    if (h->listen_fd < 0 || bind(h->listen_fd, (struct sockaddr*)&addr, alen) < 0 ||
        listen(h->listen_fd, world) < 0) {
      sy_set_error("hub: bind/listen failed for session '%s': %s", session.c_str(), strerror(errno));
      hub_destroy(h); return nullptr;
    }
    h->conn.assign(world, -1);
    for (int i = 1; i < world; ++i) {
      struct pollfd pfd = {h->listen_fd, POLLIN, 0};
      int left = (int)((deadline - now_s()) * 1000);
      if (left <= 0 || poll(&pfd, 1, left) <= 0) {
        sy_set_error("hub: timed out waiting for %d more rank(s) on session '%s'", world - i,
                     session.c_str());
        hub_destroy(h); return nullptr;
      }
      int fd = accept4(h->listen_fd, nullptr, nullptr, SOCK_CLOEXEC);
      int r = -1;
      if (fd < 0 || full_read(fd, &r, sizeof r) < 0 || r <= 0 || r >= world || h->conn[r] != -1) {
        sy_set_error("hub: bad hello on session '%s'", session.c_str());
        if (fd >= 0) close(fd);
        hub_destroy(h); return nullptr;
      }
      h->conn[r] = fd;
    }
  } else {
    h->conn.assign(1, -1);
    for (;;) {
      int fd = socket(AF_UNIX, SOCK_STREAM | SOCK_CLOEXEC, 0);
      if (fd >= 0 && connect(fd, (struct sockaddr*)&addr, alen) == 0) {
        if (full_write(fd, &rank, sizeof rank) < 0) { close(fd); fd = -1; }
        else { h->conn[0] = fd; break; }
      }
      if (fd >= 0) close(fd);
      if (now_s() > deadline) {
        sy_set_error("hub: rank %d could not reach rank 0 on session '%s'", rank, session.c_str());
        hub_destroy(h); return nullptr;
      }
      usleep(2000);
    }
  }
  return h;
}

void hub_destroy(Hub* h) {
  if (!h) return;
  for (int fd : h->conn) if (fd >= 0) close(fd);
  if (h->listen_fd >= 0) close(h->listen_fd);
  delete h;
}

int hub_allgather(Hub* h, const void* mine, size_t len, void* all) {
  if (h->world == 1) { memcpy(all, mine, len); return 0; }
  char* out = (char*)all;
  if (h->rank == 0) {
    memcpy(out, mine, len);
    for (int r = 1; r < h->world; ++r)
      if (full_read(h->conn[r], out + r * len, len) < 0) { sy_set_error("hub: allgather read"); return -1; }
    for (int r = 1; r < h->world; ++r)
      if (full_write(h->conn[r], out, len * h->world) < 0) { sy_set_error("hub: allgather write"); return -1; }
  } else {
    if (full_write(h->conn[0], mine, len) < 0 || full_read(h->conn[0], out, len * h->world) < 0) {
      sy_set_error("hub: allgather (rank %d)", h->rank); return -1;
    }
  }
  return 0;
}

int hub_barrier(Hub* h) {
  char b = 1; std::vector<char> all(h->world);
  return hub_allgather(h, &b, 1, all.data());
}

int hub_allgather_fd(Hub* h, int myfd, int* fds) {
  if (h->world == 1) { fds[0] = dup(myfd); return 0; }
  if (h->rank == 0) {
    fds[0] = dup(myfd);
    for (int r = 1; r < h->world; ++r)
      if (recv_fd(h->conn[r], &fds[r]) < 0) { sy_set_error("hub: recv_fd from %d", r); return -1; }
    for (int r = 1; r < h->world; ++r)
      for (int j = 0; j < h->world; ++j)
        if (send_fd(h->conn[r], fds[j]) < 0) { sy_set_error("hub: send_fd to %d", r); return -1; }
  } else {
    if (send_fd(h->conn[0], myfd) < 0) { sy_set_error("hub: send_fd"); return -1; }
    for (int j = 0; j < h->world; ++j)
      if (recv_fd(h->conn[0], &fds[j]) < 0) { sy_set_error("hub: recv_fd"); return -1; }
  }
  return 0;
}

int hub_bcast_fd(Hub* h, int* fd) {
  if (h->world == 1) return 0;
  if (h->rank == 0) {
    for (int r = 1; r < h->world; ++r)
      if (send_fd(h->conn[r], *fd) < 0) { sy_set_error("hub: bcast_fd"); return -1; }
  } else {
    if (recv_fd(h->conn[0], fd) < 0) { sy_set_error("hub: bcast_fd recv"); return -1; }
  }
  return 0;
}
