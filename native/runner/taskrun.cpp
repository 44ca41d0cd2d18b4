// shipyard-taskrun — native task runner for the local B200 backend.
//
// Executes ONE task (single- or multi-instance) described by a spec file:
//   system prologue -> user prologue -> env-file dump -> coordination command (once per
//   instance) -> spawn one rank per GPU (or the application command on the master only)
//   -> watchdog (rank death => whole task fails, wall-time limit, terminate signals)
//   -> system epilogue with SHIPYARD_TASK_RESULT=success|fail -> exit with the task's code.
//
// Contract parity: /root/reference/scripts/shipyard_task_runner.sh:24-63 (prologue /
// env file / user command / epilogue / exit code) and shipyard_docker_exec_task_runner.sh:29-56
// (multi-instance coordination then application phase).  Instead of `docker exec` + `mpirun`
// over ssh, ranks are fork/exec'd directly with the rank environment (RANK, WORLD_SIZE,
// LOCAL_RANK, MASTER_*, OMPI_*/PMI_* compat, AZ_BATCH_*) and the collectives shim preloaded.
//
// Spec file: one `key<TAB>value` per line; value escapes \n \t \\ ; repeated keys form lists.
#ifndef _GNU_SOURCE
#define _GNU_SOURCE
#endif
#include <errno.h>
#include <fcntl.h>
#include <grp.h>
#include <sched.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mount.h>
#include <sys/stat.h>
#include <sys/time.h>
#include <sys/types.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

#include <map>
#include <set>
#include <string>
#include <utility>
#include <vector>

extern char** environ;

struct Spec {
  std::map<std::string, std::string> kv;
  std::map<std::string, std::vector<std::string>> lists;
  std::string get(const std::string& k, const std::string& d = "") const {
    auto it = kv.find(k); return it == kv.end() ? d : it->second;
  }
  long geti(const std::string& k, long d) const {
    auto it = kv.find(k); return it == kv.end() || it->second.empty() ? d : atol(it->second.c_str());
  }
  const std::vector<std::string>& list(const std::string& k) const {
    static const std::vector<std::string> empty; auto it = lists.find(k); return it == lists.end() ? empty : it->second;
  }
};

static std::string unescape(const std::string& s) {
  std::string o; o.reserve(s.size());
  for (size_t i = 0; i < s.size(); ++i) {
    if (s[i] == '\\' && i + 1 < s.size()) {
      char c = s[++i];
      o += c == 'n' ? '\n' : c == 't' ? '\t' : c;
    } else o += s[i];
  }
  return o;
}

static bool load_spec(const char* path, Spec& sp) {
  FILE* f = fopen(path, "r");
  if (!f) { fprintf(stderr, "taskrun: cannot open spec %s: %s\n", path, strerror(errno)); return false; }
  char* line = nullptr; size_t cap = 0; ssize_t n;
  while ((n = getline(&line, &cap, f)) > 0) {
    while (n > 0 && (line[n - 1] == '\n' || line[n - 1] == '\r')) line[--n] = 0;
    if (n == 0 || line[0] == '#') continue;
    char* tab = strchr(line, '\t');
    std::string k = tab ? std::string(line, tab - line) : std::string(line);
    std::string v = tab ? unescape(tab + 1) : "";
    sp.kv[k] = v; sp.lists[k].push_back(v);
  }
  free(line); fclose(f);
  return true;
}

static double now_s() { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + ts.tv_nsec * 1e-9; }
static double wall_s() { struct timeval tv; gettimeofday(&tv, nullptr); return tv.tv_sec + tv.tv_usec * 1e-6; }

static volatile sig_atomic_t g_term = 0;
static void on_term(int) { g_term = 1; }

static int open_append(const std::string& p) {
  int fd = open(p.c_str(), O_WRONLY | O_CREAT | O_APPEND | O_CLOEXEC, 0644);
  if (fd < 0) fprintf(stderr, "taskrun: cannot open %s: %s\n", p.c_str(), strerror(errno));
  return fd;
}


// ---------------------------------------------------------------------------------------------------------------------------
// Process sandbox: the container semantics of the reference's `docker run` option synthesis
// (/root/reference/convoy/settings.py:3919-4051 user identity / working dir / default + volume binds, :3875-3901 --rm / --shm-size /
// --name, :4374-4443 daemonised multi-instance coordination container) modelled without a container runtime:
//   * a private mount namespace per rank (root: CLONE_NEWNS; unprivileged: CLONE_NEWUSER|CLONE_NEWNS with an identity uid map)
//   * `bind src:dst[:ro]`            -> recursive bind mounts (data volumes, shared data volumes, the env file ...)
//   * `restrict_root` + `keep`       -> restrict_default_bind_mounts: the node root is replaced by an empty tmpfs in which only the
//                                       task directory (and the explicitly bound volumes) are visible
//   * `shm_bytes`                    -> private tmpfs on /dev/shm of that size (--shm-size)
//   * `private_tmp`                  -> the task's own /tmp ("writable layer"); deleted after the task when `rm` is set (--rm)
//   * `uid` / `gid`                  -> setgroups/setgid/setuid after the mounts (user_identity.specific_user, -u uid:gid)
//   * `name` + `containers_dir`      -> a registry entry <name>.json (runner pid, rank process groups) while the task runs, and
//                                       <name>.coord with the session ids of the coordination commands: what `docker run -d` leaves
//                                       behind and job release / `jobs cmi` kill by name
// Modes: "mountns" (root), "userns" (unprivileged), "none" (no namespace support: binds degrade to symlinks where the destination
// does not exist; `sandbox require` turns that into a task failure with exit code 125).
struct Bind { std::string src, dst; bool ro = false; };
struct Sandbox {
  std::string want = "off";            // off | auto | require
  std::vector<Bind> binds;
  std::string restrict_root;
  std::vector<std::string> keep;
  long uid = -1, gid = -1;
  long long shm_bytes = 0;
  std::string private_tmp;
  std::string node_root;               // stays visible when it lives under /tmp and /tmp becomes private
  std::vector<std::string> keep_tmp;   // further paths under /tmp that stay visible (the shipyard checkout the task's commands use)
  bool active() const { return want != "off"; }
};

static int mkdir_p(const std::string& path, mode_t mode = 0755) {
  std::string cur;
  for (size_t i = 0; i < path.size(); ++i) {
    cur += path[i];
    if ((path[i] == '/' && i > 0) || i + 1 == path.size()) {
      if (mkdir(cur.c_str(), mode) != 0 && errno != EEXIST) return -1;
    }
  }
  return 0;
}

static bool write_file(const char* path, const std::string& v) {
  int fd = open(path, O_WRONLY | O_CLOEXEC);
  if (fd < 0) return false;
  bool ok = write(fd, v.data(), v.size()) == (ssize_t)v.size();
  close(fd);
  return ok;
}

// try to enter a private mount namespace; returns the mode achieved
static const char* sandbox_unshare() {
  if (geteuid() == 0) {
    if (unshare(CLONE_NEWNS) == 0) return "mountns";
    return "none";
  }
  const uid_t u = getuid(); const gid_t g = getgid();
  if (unshare(CLONE_NEWUSER | CLONE_NEWNS) != 0) return "none";
  write_file("/proc/self/setgroups", "deny");
  if (!write_file("/proc/self/uid_map", std::to_string(u) + " " + std::to_string(u) + " 1\n") ||
      !write_file("/proc/self/gid_map", std::to_string(g) + " " + std::to_string(g) + " 1\n")) return "none";
  return "userns";
}

// one fork probes which mode this box allows (the answer is the same for every rank of the task)
static std::string sandbox_probe() {
  pid_t pid = fork();
  if (pid == 0) {
    const char* m = sandbox_unshare();
    if (!strcmp(m, "none")) _exit(2);
    if (mount(nullptr, "/", nullptr, MS_REC | MS_PRIVATE, nullptr) != 0) _exit(2);
    _exit(!strcmp(m, "mountns") ? 0 : 1);
  }
  int st = 0;
  if (pid < 0 || waitpid(pid, &st, 0) < 0 || !WIFEXITED(st)) return "none";
  return WEXITSTATUS(st) == 0 ? "mountns" : WEXITSTATUS(st) == 1 ? "userns" : "none";
}

static bool is_dir(const std::string& p) { struct stat st; return stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode); }

// Runs in the child, before exec.  Returns 0 or an errno-style failure (the child then exits 125).
static int sandbox_enter(const Sandbox& sb, const std::string& mode) {
  if (!sb.active()) return 0;
  if (mode == "none") {
    if (sb.want == "require") { fprintf(stderr, "taskrun: sandbox required but this box allows neither mount nor user namespaces\n"); return EPERM; }
    // degraded: a bind whose destination does not exist becomes a symlink; everything else is left as is
    for (auto& b : sb.binds) {
      struct stat st;
      if (b.src == b.dst || lstat(b.dst.c_str(), &st) == 0) continue;
      size_t slash = b.dst.rfind('/');
      if (slash != std::string::npos && slash > 0) mkdir_p(b.dst.substr(0, slash));
      if (symlink(b.src.c_str(), b.dst.c_str()) != 0)
        fprintf(stderr, "taskrun: sandbox (degraded): cannot link %s -> %s: %s\n", b.dst.c_str(), b.src.c_str(), strerror(errno));
    }
    if (!sb.private_tmp.empty()) { mkdir_p(sb.private_tmp, 01777); setenv("TMPDIR", sb.private_tmp.c_str(), 1); }
    return 0;
  }
  const char* m = sandbox_unshare();
  if (strcmp(m, mode.c_str()) != 0) { fprintf(stderr, "taskrun: sandbox: unshare failed: %s\n", strerror(errno)); return EPERM; }
  if (mount(nullptr, "/", nullptr, MS_REC | MS_PRIVATE, nullptr) != 0) { perror("taskrun: sandbox: make-rprivate /"); return errno; }
  // sources are pinned by descriptor first: restrict_root may hide them from the path namespace below
  std::vector<int> src_fd(sb.binds.size(), -1);
  for (size_t i = 0; i < sb.binds.size(); ++i) {
    const Bind& b = sb.binds[i];
    struct stat st;
    if (stat(b.src.c_str(), &st) != 0 && mkdir_p(b.src) != 0) {
      fprintf(stderr, "taskrun: sandbox: bind source %s: %s\n", b.src.c_str(), strerror(errno)); return ENOENT;
    }
    src_fd[i] = open(b.src.c_str(), O_PATH | O_CLOEXEC);
    if (src_fd[i] < 0) { fprintf(stderr, "taskrun: sandbox: open %s: %s\n", b.src.c_str(), strerror(errno)); return errno; }
  }
  if (!sb.private_tmp.empty()) {
    // the task's own /tmp.  Paths under /tmp the task still needs (the node root of a state dir under /tmp, the shipyard checkout)
    // are pinned by descriptor first and re-attached at the same place inside the private /tmp.
    std::vector<std::pair<std::string, int>> pinned;
    auto pin = [&](const std::string& p) {
      if (p.compare(0, 5, "/tmp/") != 0) return;
      for (auto& q : pinned) if (p.compare(0, q.first.size() + 1, q.first + "/") == 0 || p == q.first) return;   // inside an already pinned tree
      const int fd = open(p.c_str(), O_PATH | O_CLOEXEC);
      if (fd >= 0) pinned.emplace_back(p, fd);
    };
    pin(sb.node_root);
    for (auto& k : sb.keep_tmp) pin(k);
    mkdir_p(sb.private_tmp, 01777); chmod(sb.private_tmp.c_str(), 01777);
    const int tmp_fd = open(sb.private_tmp.c_str(), O_PATH | O_CLOEXEC);
    const std::string via = "/proc/self/fd/" + std::to_string(tmp_fd);
    if (tmp_fd < 0 || mount(via.c_str(), "/tmp", nullptr, MS_BIND, nullptr) != 0)
      fprintf(stderr, "taskrun: sandbox: private /tmp: %s (keeping the host's)\n", strerror(errno));
    else
      for (auto& q : pinned) {
        const std::string rvia = "/proc/self/fd/" + std::to_string(q.second);
        if (mkdir_p(q.first) != 0 || mount(rvia.c_str(), q.first.c_str(), nullptr, MS_BIND | MS_REC, nullptr) != 0) {
          fprintf(stderr, "taskrun: sandbox: re-attach %s under the private /tmp: %s\n", q.first.c_str(), strerror(errno)); return errno;
        }
      }
    if (tmp_fd >= 0) close(tmp_fd);
    for (auto& q : pinned) close(q.second);
  }
  if (!sb.restrict_root.empty() && is_dir(sb.restrict_root)) {
    char stage[] = "/tmp/.sy-sbx-XXXXXX";
    if (!mkdtemp(stage)) { perror("taskrun: sandbox: mkdtemp"); return errno; }
    if (mount("tmpfs", stage, "tmpfs", MS_NOSUID | MS_NODEV, "mode=0755,size=16m") != 0) { perror("taskrun: sandbox: tmpfs"); return errno; }
    const std::string root = sb.restrict_root.back() == '/' ? sb.restrict_root.substr(0, sb.restrict_root.size() - 1) : sb.restrict_root;
    for (auto& k : sb.keep) {
      if (k.compare(0, root.size() + 1, root + "/") != 0) continue;
      const std::string dst = std::string(stage) + k.substr(root.size());
      if (mkdir_p(dst) != 0 || mount(k.c_str(), dst.c_str(), nullptr, MS_BIND | MS_REC, nullptr) != 0) {
        fprintf(stderr, "taskrun: sandbox: keep %s: %s\n", k.c_str(), strerror(errno)); return errno;
      }
    }
    if (mount(stage, root.c_str(), nullptr, MS_BIND | MS_REC, nullptr) != 0) { perror("taskrun: sandbox: restrict node root"); return errno; }
    umount2(stage, MNT_DETACH);
    rmdir(stage);
  }
  for (size_t i = 0; i < sb.binds.size(); ++i) {
    const Bind& b = sb.binds[i];
    struct stat st;
    const bool src_is_dir = fstat(src_fd[i], &st) == 0 && S_ISDIR(st.st_mode);
    if (stat(b.dst.c_str(), &st) != 0) {
      int rc = 0;
      if (src_is_dir) rc = mkdir_p(b.dst);
      else {
        size_t slash = b.dst.rfind('/');
        if (slash != std::string::npos && slash > 0) rc = mkdir_p(b.dst.substr(0, slash));
        int fd = open(b.dst.c_str(), O_WRONLY | O_CREAT | O_CLOEXEC, 0644);
        if (fd >= 0) close(fd); else rc = -1;
      }
      if (rc != 0) { fprintf(stderr, "taskrun: sandbox: cannot create mount point %s: %s\n", b.dst.c_str(), strerror(errno)); return errno; }
    }
    const std::string via = "/proc/self/fd/" + std::to_string(src_fd[i]);
    if (mount(via.c_str(), b.dst.c_str(), nullptr, MS_BIND | MS_REC, nullptr) != 0) {
      fprintf(stderr, "taskrun: sandbox: bind %s -> %s: %s\n", b.src.c_str(), b.dst.c_str(), strerror(errno)); return errno;
    }
    if (b.ro && mount(nullptr, b.dst.c_str(), nullptr, MS_BIND | MS_REMOUNT | MS_RDONLY | MS_REC, nullptr) != 0) {
      fprintf(stderr, "taskrun: sandbox: read-only remount of %s: %s\n", b.dst.c_str(), strerror(errno)); return errno;
    }
    close(src_fd[i]);
  }
  if (sb.shm_bytes > 0) {
    const std::string opt = "mode=1777,size=" + std::to_string(sb.shm_bytes);
    if (mount("tmpfs", "/dev/shm", "tmpfs", MS_NOSUID | MS_NODEV, opt.c_str()) != 0)
      fprintf(stderr, "taskrun: sandbox: --shm-size tmpfs on /dev/shm: %s (keeping the host's)\n", strerror(errno));
  }
  if (sb.gid >= 0 || sb.uid >= 0) {
    if (mode == "mountns") {
      gid_t g = (gid_t)(sb.gid >= 0 ? sb.gid : getgid());
      if (setgroups(1, &g) != 0 || setgid(g) != 0) { perror("taskrun: sandbox: setgid"); return errno; }
      if (sb.uid >= 0 && setuid((uid_t)sb.uid) != 0) { perror("taskrun: sandbox: setuid"); return errno; }
    } else if ((sb.uid >= 0 && (uid_t)sb.uid != getuid()) || (sb.gid >= 0 && (gid_t)sb.gid != getgid())) {
      fprintf(stderr, "taskrun: sandbox: user_identity %ld:%ld needs root; running as %d:%d\n", sb.uid, sb.gid, (int)getuid(), (int)getgid());
      if (sb.want == "require") return EPERM;
    }
  }
  return 0;
}

static void rm_rf(const std::string& path) {
  if (path.empty() || path[0] != '/' || path == "/") return;           // only absolute paths below the task directory are ever passed
  pid_t pid = fork();
  if (pid == 0) { execl("/bin/rm", "rm", "-rf", "--one-file-system", "--", path.c_str(), (char*)nullptr); _exit(127); }
  int st; if (pid > 0) waitpid(pid, &st, 0);
}

// run `cmd` through the shell with extra env; stdout/stderr to the given fds; returns exit code
static int run_shell(const std::string& shell, const std::string& cmd, const std::vector<std::string>& extra_env,
                     int out_fd, int err_fd, const std::string& cwd, const Sandbox* sb = nullptr, const std::string& sb_mode = "none",
                     pid_t* session_out = nullptr) {
  if (cmd.empty()) return 0;
  pid_t pid = fork();
  if (pid < 0) return 127;
  if (pid == 0) {
    if (session_out) setsid();           // whatever the command leaves running stays findable by session id
    if (out_fd >= 0) dup2(out_fd, 1);
    if (err_fd >= 0) dup2(err_fd, 2);
    if (sb && sandbox_enter(*sb, sb_mode) != 0) _exit(125);
    if (!cwd.empty() && chdir(cwd.c_str()) != 0) _exit(126);
    for (auto& e : extra_env) putenv(strdup(e.c_str()));
    if (out_fd >= 0) dup2(out_fd, 1);
    if (err_fd >= 0) dup2(err_fd, 2);
    execl(shell.c_str(), shell.c_str(), "-c", cmd.c_str(), (char*)nullptr);
    _exit(127);
  }
  if (session_out) *session_out = pid;
  int st = 0;
  while (waitpid(pid, &st, 0) < 0 && errno == EINTR) { if (g_term) kill(pid, SIGTERM); }
  return WIFEXITED(st) ? WEXITSTATUS(st) : 128 + WTERMSIG(st);
}

static std::string json_escape(const std::string& s) {
  std::string o;
  for (char c : s) {
    if (c == '"' || c == '\\') { o += '\\'; o += c; }
    else if (c == '\n') o += "\\n";
    else if ((unsigned char)c < 0x20) { char b[8]; snprintf(b, sizeof b, "\\u%04x", c); o += b; }
    else o += c;
  }
  return o;
}

// Environment-variable contract of the reference's task runner script (/root/reference/scripts/shipyard_task_runner.sh:24-63), so
// this binary can stand in for it: SHIPYARD_SYSTEM_PROLOGUE_CMD -> SHIPYARD_USER_PROLOGUE_CMD (a failing prologue aborts with its
// code: the script runs under `set -e`) -> environment minus SHIPYARD_ENV_EXCLUDE (grep -E pattern) written to SHIPYARD_ENV_FILE ->
// `$SHIPYARD_RUNTIME $SHIPYARD_RUNTIME_CMD $SHIPYARD_RUNTIME_CMD_OPTS $SHIPYARD_CONTAINER_IMAGE_NAME $SHIPYARD_USER_CMD` (or the bare
// user command) -> SHIPYARD_SYSTEM_EPILOGUE_CMD with SHIPYARD_TASK_RESULT=success|fail -> exit with the task's code.
static int run_env_contract() {
  auto env = [](const char* k) { const char* v = getenv(k); return std::string(v ? v : ""); };
  const std::string shell = "/bin/bash";
  int rc = run_shell(shell, env("SHIPYARD_SYSTEM_PROLOGUE_CMD"), {}, -1, -1, "");
  if (rc != 0) return rc;
  rc = run_shell(shell, env("SHIPYARD_USER_PROLOGUE_CMD"), {}, -1, -1, "");
  if (rc != 0) return rc;
  if (!env("SHIPYARD_ENV_FILE").empty()) {
    const std::string dump = env("SHIPYARD_ENV_EXCLUDE").empty() ? "env > \"$SHIPYARD_ENV_FILE\""
                                                                 : "env | grep -vE \"$SHIPYARD_ENV_EXCLUDE\" > \"$SHIPYARD_ENV_FILE\"";
    rc = run_shell(shell, dump, {}, -1, -1, "");
    if (rc != 0 && env("SHIPYARD_ENV_EXCLUDE").empty()) return rc;      // (grep exits 1 when every line was excluded: not an error)
  }
  const std::string task = env("SHIPYARD_RUNTIME").empty()
      ? "eval \"$SHIPYARD_USER_CMD\""
      : "SHIPYARD_RUNTIME_CMD_OPTS=$(eval echo \"${SHIPYARD_RUNTIME_CMD_OPTS}\"); "
        "eval \"$SHIPYARD_RUNTIME $SHIPYARD_RUNTIME_CMD $SHIPYARD_RUNTIME_CMD_OPTS $SHIPYARD_CONTAINER_IMAGE_NAME $SHIPYARD_USER_CMD\"";
  const int task_rc = run_shell(shell, task, {}, -1, -1, "");
  if (!env("SHIPYARD_SYSTEM_EPILOGUE_CMD").empty())
    run_shell(shell, "eval \"$SHIPYARD_SYSTEM_EPILOGUE_CMD\"", {std::string("SHIPYARD_TASK_RESULT=") + (task_rc == 0 ? "success" : "fail")}, -1, -1, "");
  return task_rc;
}

int main(int argc, char** argv) {
  const char* spec_path = nullptr;
  for (int i = 1; i < argc; ++i) {
    if (!strcmp(argv[i], "--spec") && i + 1 < argc) spec_path = argv[++i];
    else if (!strcmp(argv[i], "--version")) { puts("shipyard-taskrun 0.1"); return 0; }
  }
  if (!spec_path && getenv("SHIPYARD_USER_CMD")) return run_env_contract();
  if (!spec_path) { fprintf(stderr, "usage: shipyard-taskrun --spec <file>   (or the SHIPYARD_USER_CMD environment contract)\n"); return 2; }
  Spec sp;
  if (!load_spec(spec_path, sp)) return 2;

  const std::string shell = sp.get("shell", "/bin/bash");
  const std::string workdir = sp.get("workdir", ".");
  const std::string taskdir = sp.get("taskdir", workdir);
  const std::string out_path = sp.get("stdout", taskdir + "/stdout.txt");
  const std::string err_path = sp.get("stderr", taskdir + "/stderr.txt");
  const long world_cfg = sp.geti("world", 1);
  const long rpi = sp.geti("ranks_per_instance", 1);
  const long ninst = sp.geti("num_instances", 1);
  const bool master_only = sp.geti("master_only", 0) != 0;
  const long wall_time_s = sp.geti("wall_time_s", 0);
  const std::string preload = sp.get("preload");
  const std::string session = sp.get("session", "task");
  const std::string master_port = sp.get("master_port", "29400");
  const auto& gpus = sp.list("gpu");

  Sandbox sb;
  sb.want = sp.get("sandbox", "off");
  for (auto& b : sp.list("bind")) {          // docker syntax: src:dst[:ro|rw]
    Bind bd; size_t c1 = b.find(':');
    if (c1 == std::string::npos) { bd.src = bd.dst = b; }
    else {
      bd.src = b.substr(0, c1); size_t c2 = b.find(':', c1 + 1);
      bd.dst = b.substr(c1 + 1, c2 == std::string::npos ? std::string::npos : c2 - c1 - 1);
      if (c2 != std::string::npos) bd.ro = b.substr(c2 + 1).find("ro") != std::string::npos;
    }
    if (!bd.src.empty() && !bd.dst.empty()) sb.binds.push_back(bd);
  }
  sb.restrict_root = sp.get("restrict_root");
  sb.keep = sp.list("keep");
  sb.uid = sp.geti("uid", -1); sb.gid = sp.geti("gid", -1);
  sb.shm_bytes = atoll(sp.get("shm_bytes", "0").c_str());
  sb.private_tmp = sp.get("private_tmp");
  sb.node_root = sp.get("node_root");
  sb.keep_tmp = sp.list("keep_tmp");
  const bool rm_after_exit = sp.geti("rm", 0) != 0;
  const std::string ctr_name = sp.get("name"), ctr_dir = sp.get("containers_dir"), ctr_scratch = sp.get("container_scratch");
  const std::string sb_mode = sb.active() ? sandbox_probe() : "none";
  if (sb.active()) setenv("SHIPYARD_SANDBOX", sb_mode.c_str(), 1);

  struct sigaction sa = {};
  sa.sa_handler = on_term;
  sigaction(SIGTERM, &sa, nullptr);
  sigaction(SIGINT, &sa, nullptr);
  signal(SIGPIPE, SIG_IGN);

  mkdir(workdir.c_str(), 0755);
  if (sb.active() && sb.uid >= 0 && geteuid() == 0) {
    // user_identity.specific_user: the task user owns its working and task directories (as on a Batch node)
    const gid_t g = (gid_t)(sb.gid >= 0 ? sb.gid : sb.uid);
    if (chown(workdir.c_str(), (uid_t)sb.uid, g) != 0 || chown(taskdir.c_str(), (uid_t)sb.uid, g) != 0)
      fprintf(stderr, "taskrun: chown of the task directories to %ld:%ld failed: %s\n", sb.uid, sb.gid, strerror(errno));
  }
  for (auto& e : sp.list("env")) putenv(strdup(e.c_str()));
  int out_fd = open_append(out_path), err_fd = open_append(err_path);
  const double t_start = wall_s();

  int rc = 0;
  std::string phase = "prologue";
  // ---- prologues -----------------------------------------------------------------
  rc = run_shell(shell, sp.get("system_prologue"), {}, out_fd, err_fd, workdir);
  if (rc == 0) rc = run_shell(shell, sp.get("user_prologue"), {}, out_fd, err_fd, workdir);

  // ---- environment file (what a container would receive via --env-file) ----------
  const std::string env_file = sp.get("env_file");
  if (rc == 0 && !env_file.empty()) {
    std::set<std::string> excl(sp.list("env_exclude").begin(), sp.list("env_exclude").end());
    FILE* ef = fopen(env_file.c_str(), "w");
    if (ef) {
      for (char** e = environ; *e; ++e) {
        const char* eq = strchr(*e, '=');
        if (!eq) continue;
        std::string name(*e, eq - *e);
        if (excl.count(name) || strchr(*e, '\n')) continue;
        fprintf(ef, "%s\n", *e);
      }
      fclose(ef);
    }
  }

  // ---- coordination phase: once per instance, must not block ----------------------
  const std::string coord = sp.get("coordination_cmd");
  if (rc == 0 && !coord.empty()) {
    phase = "coordination";
    for (long i = 0; i < ninst && rc == 0; ++i) {
      std::vector<std::string> env = {"SHIPYARD_INSTANCE=" + std::to_string(i),
                                      std::string("AZ_BATCH_IS_CURRENT_NODE_MASTER=") + (i == 0 ? "true" : "false")};
      pid_t sid = 0;
      rc = run_shell(shell, coord, env, out_fd, err_fd, workdir, &sb, sb_mode, &sid);
      if (sid > 0 && !ctr_dir.empty() && !ctr_name.empty()) {
        // like `docker run -d`: whatever the coordination command daemonised outlives this task; job release / `jobs cmi` find it here
        mkdir_p(ctr_dir);
        FILE* cf = fopen((ctr_dir + "/" + ctr_name + ".coord").c_str(), "a");
        if (cf) { fprintf(cf, "%d\n", (int)sid); fclose(cf); }
      }
    }
  }

  // ---- application phase ----------------------------------------------------------
  std::vector<pid_t> pids;
  std::vector<int> codes;
  bool timed_out = false, terminated = false;
  const std::string user_cmd = sp.get("user_cmd");
  const long world = master_only ? 1 : world_cfg;
  if (rc == 0 && !user_cmd.empty()) {
    phase = "application";
    // fault injection for tests: kill_rank:<k>:after_ms:<t>
    long fi_rank = -1; double fi_after = 0;
    if (const char* fi = getenv("SHIPYARD_FAULT_INJECT")) {
      long k = 0, t = 0;
      if (sscanf(fi, "kill_rank:%ld:after_ms:%ld", &k, &t) == 2) { fi_rank = k; fi_after = t * 1e-3; }
    }
    pids.assign(world, -1); codes.assign(world, -1);
    for (long r = 0; r < world; ++r) {
      pid_t pid = fork();
      if (pid < 0) { rc = 127; break; }
      if (pid == 0) {
        setpgid(0, 0);
        const long local = r % rpi, inst = r / rpi;
        auto set = [](const std::string& k, const std::string& v) { setenv(k.c_str(), v.c_str(), 1); };
        set("RANK", std::to_string(r)); set("WORLD_SIZE", std::to_string(world));
        set("LOCAL_RANK", std::to_string(world_cfg == world ? r : 0));   // one box: local == global
        set("LOCAL_WORLD_SIZE", std::to_string(world));
        set("GROUP_RANK", "0"); set("SHIPYARD_INSTANCE", std::to_string(inst));
        set("SHIPYARD_INSTANCE_LOCAL_RANK", std::to_string(local));
        set("MASTER_ADDR", "127.0.0.1"); set("MASTER_PORT", master_port);
        set("SHIPYARD_RANK", std::to_string(r)); set("SHIPYARD_WORLD_SIZE", std::to_string(world));
        set("SHIPYARD_COLL_SESSION", session);
        set("OMPI_COMM_WORLD_RANK", std::to_string(r)); set("OMPI_COMM_WORLD_SIZE", std::to_string(world));
        set("OMPI_COMM_WORLD_LOCAL_RANK", std::to_string(r)); set("OMPI_COMM_WORLD_LOCAL_SIZE", std::to_string(world));
        set("PMI_RANK", std::to_string(r)); set("PMI_SIZE", std::to_string(world));
        set("AZ_BATCH_IS_CURRENT_NODE_MASTER", inst == 0 ? "true" : "false");
        if ((size_t)r < gpus.size() && !master_only) set("SHIPYARD_GPU", gpus[r]);
        if (!master_only && !gpus.empty() && gpus[0] != "-1") set("CUDA_DEVICE_ORDER", "PCI_BUS_ID");
        if (!preload.empty()) {
          const char* old = getenv("LD_PRELOAD");
          set("LD_PRELOAD", old && *old ? preload + ":" + old : preload);
        }
        int o = out_fd, e = err_fd;
        if (r > 0) {
          o = open_append(taskdir + "/stdout.r" + std::to_string(r) + ".txt");
          e = open_append(taskdir + "/stderr.r" + std::to_string(r) + ".txt");
        }
        if (o >= 0) dup2(o, 1);
        if (e >= 0) dup2(e, 2);
        if (sandbox_enter(sb, sb_mode) != 0) _exit(125);
        if (chdir(workdir.c_str()) != 0) _exit(126);
        execl(shell.c_str(), shell.c_str(), "-c", user_cmd.c_str(), (char*)nullptr);
        _exit(127);
      }
      setpgid(pid, pid);
      pids[r] = pid;
    }
    {
      // rank process groups: `jobs tasks term --force` / `jobs del` kill these if the runner itself has to be killed
      FILE* pf = fopen((taskdir + "/ranks.pid").c_str(), "w");
      if (pf) { for (auto p : pids) if (p > 0) fprintf(pf, "%d\n", (int)p); fclose(pf); }
      if (!ctr_dir.empty() && !ctr_name.empty()) {
        mkdir_p(ctr_dir);
        FILE* cf = fopen((ctr_dir + "/" + ctr_name + ".json").c_str(), "w");
        if (cf) {
          fprintf(cf, "{\"name\": \"%s\", \"runner_pid\": %d, \"sandbox\": \"%s\", \"taskdir\": \"%s\", \"rank_pgids\": [",
                  json_escape(ctr_name).c_str(), (int)getpid(), sb_mode.c_str(), json_escape(taskdir).c_str());
          bool first = true;
          for (auto p : pids) if (p > 0) { fprintf(cf, "%s%d", first ? "" : ", ", (int)p); first = false; }
          fprintf(cf, "]}\n"); fclose(cf);
        }
      }
    }
    // ---- watchdog loop --------------------------------------------------------------
    const double t0 = now_s();
    long alive = 0;
    for (auto p : pids) if (p > 0) ++alive;
    double kill_deadline = 0;       // when set: escalate SIGTERM -> SIGKILL
    bool failing = rc != 0;
    auto signal_all = [&](int sig) { for (size_t i = 0; i < pids.size(); ++i) if (pids[i] > 0 && codes[i] < 0) kill(-pids[i], sig); };
    if (failing) { signal_all(SIGTERM); kill_deadline = now_s() + 5; }
    const std::string hb = sp.get("heartbeat");
    double next_hb = 0;
    while (alive > 0) {
      int st = 0;
      pid_t w = waitpid(-1, &st, WNOHANG);
      if (w > 0) {
        for (size_t i = 0; i < pids.size(); ++i) {
          if (pids[i] != w) continue;
          codes[i] = WIFEXITED(st) ? WEXITSTATUS(st) : 128 + WTERMSIG(st);
          --alive;
          if (codes[i] != 0 && !failing) {
            // one rank died: the whole multi-instance task fails; reap the rest instead of hanging
            failing = true; rc = codes[i];
            fprintf(stderr, "taskrun: rank %zu exited with %d; terminating %ld remaining rank(s)\n", i, codes[i], alive);
            signal_all(SIGTERM); kill_deadline = now_s() + 5;
          }
        }
        continue;
      }
      const double t = now_s();
      if (g_term && !terminated) {
        terminated = true; failing = true; if (rc == 0) rc = 143;
        signal_all(SIGTERM); kill_deadline = t + 5;
      }
      if (wall_time_s > 0 && t - t0 > (double)wall_time_s && !timed_out) {
        timed_out = true; failing = true; rc = 124;
        fprintf(stderr, "taskrun: wall time limit of %ld s exceeded\n", wall_time_s);
        signal_all(SIGTERM); kill_deadline = t + 5;
      }
      if (fi_rank >= 0 && fi_rank < (long)pids.size() && t - t0 > fi_after && codes[fi_rank] < 0) {
        kill(-pids[fi_rank], SIGKILL); fi_rank = -1;
      }
      if (kill_deadline > 0 && t > kill_deadline) { signal_all(SIGKILL); kill_deadline = t + 60; }
      if (!hb.empty() && t > next_hb) {
        int fd = open(hb.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
        if (fd >= 0) { dprintf(fd, "%.3f\n", wall_s()); close(fd); }
        next_hb = t + 2.0;
      }
      usleep(10000);
    }
    if (rc == 0) for (int c : codes) if (c > 0) { rc = c; break; }
  }

  // ---- epilogue: always runs, sees the result -------------------------------------
  const std::string result = rc == 0 ? "success" : "fail";
  int erc = run_shell(shell, sp.get("system_epilogue"), {"SHIPYARD_TASK_RESULT=" + result}, out_fd, err_fd, workdir);
  if (erc != 0) fprintf(stderr, "taskrun: system epilogue exited with %d\n", erc);

  unlink((taskdir + "/ranks.pid").c_str());
  if (!ctr_dir.empty() && !ctr_name.empty()) unlink((ctr_dir + "/" + ctr_name + ".json").c_str());
  if (rm_after_exit && !ctr_scratch.empty()) rm_rf(ctr_scratch);      // --rm: the "container" (private /tmp, anonymous volumes) goes away

  const std::string rf = sp.get("result_file");
  if (!rf.empty()) {
    std::string tmp = rf + ".tmp";
    FILE* f = fopen(tmp.c_str(), "w");
    if (f) {
      fprintf(f, "{\"exit_code\": %d, \"result\": \"%s\", \"phase\": \"%s\", \"start\": %.3f, \"end\": %.3f, "
                 "\"timed_out\": %s, \"terminated\": %s, \"world\": %ld, \"sandbox\": \"%s\", \"rank_exit_codes\": [",
              rc, result.c_str(), json_escape(phase).c_str(), t_start, wall_s(), timed_out ? "true" : "false",
              terminated ? "true" : "false", world, sb.active() ? sb_mode.c_str() : "off");
      for (size_t i = 0; i < codes.size(); ++i) fprintf(f, "%s%d", i ? ", " : "", codes[i]);
      fprintf(f, "]}\n");
      fclose(f);
      rename(tmp.c_str(), rf.c_str());
    }
  }
  return rc;
}
