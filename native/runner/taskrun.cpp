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
#include <errno.h>
#include <fcntl.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <sys/time.h>
#include <sys/types.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

#include <map>
#include <set>
#include <string>
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

// run `cmd` through the shell with extra env; stdout/stderr to the given fds; returns exit code
static int run_shell(const std::string& shell, const std::string& cmd, const std::vector<std::string>& extra_env,
                     int out_fd, int err_fd, const std::string& cwd) {
  if (cmd.empty()) return 0;
  pid_t pid = fork();
  if (pid < 0) return 127;
  if (pid == 0) {
    if (!cwd.empty() && chdir(cwd.c_str()) != 0) _exit(126);
    for (auto& e : extra_env) putenv(strdup(e.c_str()));
    if (out_fd >= 0) dup2(out_fd, 1);
    if (err_fd >= 0) dup2(err_fd, 2);
    execl(shell.c_str(), shell.c_str(), "-c", cmd.c_str(), (char*)nullptr);
    _exit(127);
  }
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

  struct sigaction sa = {};
  sa.sa_handler = on_term;
  sigaction(SIGTERM, &sa, nullptr);
  sigaction(SIGINT, &sa, nullptr);
  signal(SIGPIPE, SIG_IGN);

  mkdir(workdir.c_str(), 0755);
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
      rc = run_shell(shell, coord, env, out_fd, err_fd, workdir);
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
        if (chdir(workdir.c_str()) != 0) _exit(126);
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
        execl(shell.c_str(), shell.c_str(), "-c", user_cmd.c_str(), (char*)nullptr);
        _exit(127);
      }
      setpgid(pid, pid);
      pids[r] = pid;
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

  const std::string rf = sp.get("result_file");
  if (!rf.empty()) {
    std::string tmp = rf + ".tmp";
    FILE* f = fopen(tmp.c_str(), "w");
    if (f) {
      fprintf(f, "{\"exit_code\": %d, \"result\": \"%s\", \"phase\": \"%s\", \"start\": %.3f, \"end\": %.3f, "
                 "\"timed_out\": %s, \"terminated\": %s, \"world\": %ld, \"rank_exit_codes\": [",
              rc, result.c_str(), json_escape(phase).c_str(), t_start, wall_s(), timed_out ? "true" : "false",
              terminated ? "true" : "false", world);
      for (size_t i = 0; i < codes.size(); ++i) fprintf(f, "%s%d", i ? ", " : "", codes[i]);
      fprintf(f, "]}\n");
      fclose(f);
      rename(tmp.c_str(), rf.c_str());
    }
  }
  return rc;
}
