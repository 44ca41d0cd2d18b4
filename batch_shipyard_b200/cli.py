"""``shipyard`` command line: same 13 groups / verbs / shared options as the reference CLI.

Reference: /root/reference/shipyard.py (click tree :1001-3127; shared option
decorators with ``SHIPYARD_*`` env fallbacks :582-998; per-domain context
initialisation :289-575).  Groups: account cert data diag fed fs jobs keyvault
misc monitor pool slurm storage.  Output: ``--raw`` prints JSON, otherwise a
YAML-ish readable dump.  Exit codes: ``pool exists`` returns 1 when absent,
configuration errors exit 1.
"""
from __future__ import annotations

import functools
import json
import os
import sys
from typing import Any, Callable, Optional

import click
import yaml

from . import __version__, fleet
from .config import loader
from .config.schema import ConfigType
from .utils import util

_CONTEXT_SETTINGS = dict(help_option_names=["-h", "--help"])


class CliContext:
    def __init__(self):
        self.verbose = False
        self.yes = False
        self.raw = False
        self.show_config = False
        self.configdir: Optional[str] = None
        self.log_file: Optional[str] = None
        self.paths: dict = {}
        self.ctx: Optional[fleet.Context] = None
        # the reference's cloud flags: recorded; only keyvault-credentials-secret-id has a local meaning (see init)
        self.compat: dict = {}

    def init(self, required: tuple = (), skip: tuple = ()) -> fleet.Context:
        util.setup_logger("shipyard", self.verbose, logfile=self.log_file)
        # --keyvault-credentials-secret-id: the credentials section comes from the (local) key vault instead of credentials.yaml
        # (/root/reference/shipyard.py:441-575 fetches it from Azure KeyVault at the same point)
        vault_id = self.compat.get("keyvault-credentials-secret-id")
        if vault_id:
            required = tuple(k for k in required if k != ConfigType.Credentials)
            skip = tuple(skip) + (ConfigType.Credentials,)
        try:
            config = loader.load_configs(self.paths, self.configdir, required=required, skip=skip, verbose=self.verbose,
                                         auto_confirm=self.yes, raw=self.raw)
            if vault_id:
                from . import keyvault
                try:
                    creds = keyvault.fetch_credentials(fleet.resolve_state_dir(config), vault_id)
                except KeyError as e:
                    raise loader.ConfigError(str(e).strip("'\"")) from e
                config = util.merge_dict(creds if "credentials" in (creds or {}) else {"credentials": creds or {}}, config)
        except loader.ConfigError as e:
            click.echo(f"ERROR: {e}", err=True)
            sys.exit(1)
        if self.show_config:
            click.echo(loader.dump_config(config), err=True)
        state_dir = fleet.resolve_state_dir(config)
        self.ctx = fleet.Context(config=config, state_dir=state_dir, raw=self.raw, yes=self.yes, verbose=self.verbose,
                                 inline_agent=bool(os.environ.get("SHIPYARD_INLINE_AGENT")))
        return self.ctx


pass_cli_context = click.make_pass_decorator(CliContext, ensure=True)


def _setter(name: str, kind: Optional[ConfigType] = None):
    def cb(ctx, param, value):
        c = ctx.ensure_object(CliContext)
        if kind is not None:
            if value:
                c.paths[kind] = value
        elif name.startswith("compat:"):
            if value is not None:
                c.compat[name[7:]] = value
        elif value is not None and value is not False or name in ("verbose", "yes", "raw", "show_config"):
            if value:
                setattr(c, name, value)
        return value
    return cb


def common_options(f):
    opts = [
        click.option("-y", "--yes", is_flag=True, expose_value=False, envvar="SHIPYARD_YES", callback=_setter("yes"),
                     help="Assume yes for all confirmation prompts"),
        click.option("--show-config", is_flag=True, expose_value=False, envvar="SHIPYARD_SHOW_CONFIG", callback=_setter("show_config"),
                     help="Show the merged configuration"),
        click.option("-v", "--verbose", is_flag=True, expose_value=False, envvar="SHIPYARD_VERBOSE", callback=_setter("verbose"),
                     help="Verbose output"),
        click.option("--raw", is_flag=True, expose_value=False, envvar="SHIPYARD_RAW", callback=_setter("raw"),
                     help="Output data as returned by the backend, as JSON"),
        click.option("--log-file", expose_value=False, envvar="SHIPYARD_LOG_FILE", callback=_setter("log_file"),
                     help="Write log messages to this file instead of stderr"),
        click.option("--configdir", expose_value=False, envvar="SHIPYARD_CONFIGDIR", callback=_setter("configdir"),
                     help="Configuration directory holding all configuration files"),
        click.option("--credentials", expose_value=False, envvar="SHIPYARD_CREDENTIALS_CONF",
                     callback=_setter("", ConfigType.Credentials), help="Credentials config file"),
        click.option("--config", expose_value=False, envvar="SHIPYARD_CONFIG_CONF", callback=_setter("", ConfigType.Global),
                     help="Global config file"),
    ]
    for o in reversed(opts):
        f = o(f)
    return f


def _file_option(flag: str, env: str, kind: ConfigType, text: str):
    return click.option(flag, expose_value=False, envvar=env, callback=_setter("", kind), help=text)


pool_option = _file_option("--pool", "SHIPYARD_POOL_CONF", ConfigType.Pool, "Pool config file")
jobs_option = _file_option("--jobs", "SHIPYARD_JOBS_CONF", ConfigType.Jobs, "Jobs config file")
fs_option = _file_option("--fs", "SHIPYARD_FS_CONF", ConfigType.RemoteFS, "RemoteFS config file")
monitor_option = _file_option("--monitor", "SHIPYARD_MONITOR_CONF", ConfigType.Monitor, "Monitoring config file")
federation_option = _file_option("--federation", "SHIPYARD_FEDERATION_CONF", ConfigType.Federation, "Federation config file")
slurm_option = _file_option("--slurm", "SHIPYARD_SLURM_CONF", ConfigType.Slurm, "Slurm config file")


def cloud_compat_options(f):
    """AAD / key vault / subscription flags of the reference: parsed, recorded, not needed locally."""
    names = [("--aad-authority-url", "SHIPYARD_AAD_AUTHORITY_URL"), ("--aad-directory-id", "SHIPYARD_AAD_DIRECTORY_ID"),
             ("--aad-application-id", "SHIPYARD_AAD_APPLICATION_ID"), ("--aad-auth-key", "SHIPYARD_AAD_AUTH_KEY"),
             ("--aad-user", "SHIPYARD_AAD_USER"), ("--aad-password", "SHIPYARD_AAD_PASSWORD"),
             ("--aad-cert-private-key", "SHIPYARD_AAD_CERT_PRIVATE_KEY"), ("--aad-cert-thumbprint", "SHIPYARD_AAD_CERT_THUMBPRINT"),
             ("--aad-endpoint", "SHIPYARD_AAD_ENDPOINT"), ("--keyvault-uri", "SHIPYARD_KEYVAULT_URI"),
             ("--keyvault-credentials-secret-id", "SHIPYARD_KEYVAULT_CREDENTIALS_SECRET_ID"),
             ("--subscription-id", "SHIPYARD_SUBSCRIPTION_ID")]
    for flag, env in reversed(names):
        f = click.option(flag, expose_value=False, envvar=env, hidden=True, callback=_setter("compat:" + flag[2:]))(f)
    return f


def emit(cctx: CliContext, value: Any) -> None:
    if cctx.raw:
        click.echo(json.dumps(value, indent=2, sort_keys=True, default=str))
    elif isinstance(value, (dict, list)):
        click.echo(yaml.safe_dump(json.loads(json.dumps(value, default=str)), default_flow_style=False, sort_keys=False).rstrip())
    elif value is not None:
        click.echo(str(value))


def run_action(cctx: CliContext, fn: Callable, *a, **kw) -> Any:
    try:
        out = fn(cctx.ctx, *a, **kw)
    except (KeyError, ValueError, RuntimeError, loader.ConfigError, OSError) as e:
        # every domain error of the package derives from RuntimeError / ValueError (ActionError, BackendError, FederationError,
        # RemoteFsError, JobSubmissionError, PoolCreationError, FormulaError, ...): report it like the reference does, no traceback
        msg = e.args[0] if isinstance(e, KeyError) and e.args else e
        click.echo(f"ERROR: {msg}", err=True)
        sys.exit(1)
    emit(cctx, out)
    return out


_ALL = (ConfigType.Credentials, ConfigType.Global, ConfigType.Pool, ConfigType.Jobs, ConfigType.RemoteFS,
        ConfigType.Monitor, ConfigType.Federation, ConfigType.Slurm)


def _skip_except(*keep) -> tuple:
    return tuple(k for k in _ALL if k not in keep)


BATCH = (ConfigType.Credentials, ConfigType.Global, ConfigType.Pool, ConfigType.Jobs)


@click.group(context_settings=_CONTEXT_SETTINGS)
@click.version_option(version=__version__)
@click.pass_context
def cli(ctx):
    """Batch Shipyard for one B200 box: provision the GPU pool, run containerised batch / multi-instance jobs."""
    ctx.ensure_object(CliContext)


# ---------------------------------------------------------------------------------------------- account
@cli.group()
def account():
    """Account (box) actions"""


@account.command("info")
@click.option("--name")
@click.option("--resource-group")
@common_options
@cloud_compat_options
@pass_cli_context
def account_info(c, name, resource_group):
    """Retrieve box / account information"""
    c.init(skip=_skip_except(ConfigType.Credentials, ConfigType.Global)); run_action(c, fleet.action_account_info, name, resource_group)


@account.command("list")
@click.option("--resource-group")
@common_options
@cloud_compat_options
@pass_cli_context
def account_list(c, resource_group):
    """List accounts"""
    c.init(skip=_skip_except(ConfigType.Credentials, ConfigType.Global)); run_action(c, fleet.action_account_list, resource_group)


@account.command("quota")
@click.argument("location", required=False)
@common_options
@cloud_compat_options
@pass_cli_context
def account_quota(c, location):
    """Retrieve GPU / slot quota of the box"""
    c.init(skip=_skip_except(ConfigType.Credentials, ConfigType.Global)); run_action(c, fleet.action_account_quota, location)


@account.command("images")
@click.option("--show-unrelated", is_flag=True)
@click.option("--show-unverified", is_flag=True)
@common_options
@cloud_compat_options
@pass_cli_context
def account_images(c, show_unrelated, show_unverified):
    """List image artefacts available in the local image store"""
    c.init(skip=_skip_except(ConfigType.Credentials, ConfigType.Global)); run_action(c, fleet.action_account_images, show_unrelated, show_unverified)


# ---------------------------------------------------------------------------------------------- pool
@cli.group()
def pool():
    """Pool actions"""


def _pool_cmd(name, help_, opts=(), need=(ConfigType.Pool,), keep=BATCH):
    def deco(fn):
        @pool.command(name, help=help_)
        @common_options
        @pool_option
        @jobs_option
        @cloud_compat_options
        @pass_cli_context
        @functools.wraps(fn)
        def wrapper(c, **kw):
            c.init(required=need, skip=_skip_except(*keep))
            return fn(c, **kw)
        for o in reversed(opts):
            wrapper = o(wrapper)
        return wrapper
    return deco


@_pool_cmd("add", "Add (provision) the pool on this box", [click.option("--recreate", is_flag=True, help="Recreate the pool if it exists"),
                                                            click.option("--no-wait", is_flag=True, help="Do not wait for nodes to become ready")])
def pool_add(c, recreate, no_wait):
    run_action(c, fleet.action_pool_add, recreate, no_wait)


@_pool_cmd("exists", "Check if a pool exists (exit code 1 when it does not)", [click.option("--pool-id")], need=())
def pool_exists(c, pool_id):
    try:
        ok = fleet.action_pool_exists(c.ctx, pool_id)
    except KeyError as e:
        click.echo(f"ERROR: {e}", err=True); sys.exit(1)
    emit(c, {"exists": ok} if c.raw else f"pool {'exists' if ok else 'does not exist'}")
    sys.exit(0 if ok else 1)


@_pool_cmd("list", "List all pools", need=())
def pool_list(c):
    run_action(c, fleet.action_pool_list)


@_pool_cmd("del", "Delete a pool", [click.option("--poolid"), click.option("--wait", is_flag=True)], need=())
def pool_del(c, poolid, wait):
    run_action(c, fleet.action_pool_delete, poolid, wait)


@_pool_cmd("resize", "Resize the pool to vm_count of the pool config", [click.option("--wait", is_flag=True)])
def pool_resize(c, wait):
    run_action(c, fleet.action_pool_resize, wait)


@_pool_cmd("stats", "Get pool statistics", [click.option("--poolid")], need=())
def pool_stats(c, poolid):
    run_action(c, fleet.action_pool_stats, poolid)


@_pool_cmd("ssh", "Run a command in a node's context (no remote host on a local pool)",
           [click.option("--cardinal", type=int), click.option("--nodeid"), click.option("--tty", is_flag=True), click.argument("command", nargs=-1)])
def pool_ssh(c, cardinal, nodeid, tty, command):
    run_action(c, fleet.action_pool_ssh, cardinal, nodeid, tty, command)


@_pool_cmd("rdp", "RDP to a node (not applicable locally)", [click.option("--cardinal", type=int), click.option("--no-auto", is_flag=True), click.option("--nodeid")])
def pool_rdp(c, cardinal, no_auto, nodeid):
    run_action(c, fleet.action_pool_rdp, cardinal, no_auto, nodeid)


@pool.group("autoscale")
def pool_autoscale():
    """Autoscale actions"""


@pool.group("images")
def pool_images():
    """Container image actions"""


@pool.group("user")
def pool_user():
    """Remote user actions"""


@pool.group("nodes")
def pool_nodes():
    """Compute node actions"""


def _sub_cmd(group, name, help_, opts=(), need=(ConfigType.Pool,), keep=BATCH, extra_file_opts=()):
    def deco(fn):
        @group.command(name, help=help_)
        @common_options
        @pool_option
        @jobs_option
        @cloud_compat_options
        @pass_cli_context
        @functools.wraps(fn)
        def wrapper(c, **kw):
            c.init(required=need, skip=_skip_except(*keep))
            return fn(c, **kw)
        for o in reversed(tuple(extra_file_opts) + tuple(opts)):
            wrapper = o(wrapper)
        return wrapper
    return deco


@_sub_cmd(pool_autoscale, "enable", "Enable autoscale on the pool")
def autoscale_enable(c):
    run_action(c, fleet.action_pool_autoscale_enable)


@_sub_cmd(pool_autoscale, "disable", "Disable autoscale on the pool")
def autoscale_disable(c):
    run_action(c, fleet.action_pool_autoscale_disable)


@_sub_cmd(pool_autoscale, "evaluate", "Evaluate the autoscale formula without applying it")
def autoscale_evaluate(c):
    run_action(c, fleet.action_pool_autoscale_evaluate)


@_sub_cmd(pool_autoscale, "lastexec", "Result of the last autoscale evaluation")
def autoscale_lastexec(c):
    run_action(c, fleet.action_pool_autoscale_lastexec)


@_sub_cmd(pool_images, "list", "List images pre-loaded on the pool")
def images_list(c):
    run_action(c, fleet.action_pool_images_list)


@_sub_cmd(pool_images, "update", "Update (re-load) container images on all nodes",
          [click.option("--docker-image"), click.option("--docker-image-digest"), click.option("--singularity-image"), click.option("--ssh", is_flag=True)])
def images_update(c, docker_image, docker_image_digest, singularity_image, ssh):
    run_action(c, fleet.action_pool_images_update, docker_image, docker_image_digest, singularity_image, ssh)


@_sub_cmd(pool_user, "add", "Add a remote user (records a key pair)")
def user_add(c):
    run_action(c, fleet.action_pool_user_add)


@_sub_cmd(pool_user, "del", "Delete the remote user")
def user_del(c):
    run_action(c, fleet.action_pool_user_del)


@_sub_cmd(pool_nodes, "list", "List nodes (GPUs) in the pool", [click.option("--start-task-failed", is_flag=True), click.option("--unusable", is_flag=True)])
def nodes_list(c, start_task_failed, unusable):
    run_action(c, fleet.action_pool_nodes_list, start_task_failed, unusable)


@_sub_cmd(pool_nodes, "count", "Node counts by state", [click.option("--poolid")], need=())
def nodes_count(c, poolid):
    run_action(c, fleet.action_pool_nodes_count, poolid)


@_sub_cmd(pool_nodes, "grls", "Get remote login settings for all nodes", [click.option("--no-generate-tunnel-script", is_flag=True)])
def nodes_grls(c, no_generate_tunnel_script):
    run_action(c, fleet.action_pool_nodes_grls, no_generate_tunnel_script)


@_sub_cmd(pool_nodes, "del", "Delete nodes from the pool",
          [click.option("--all-start-task-failed", is_flag=True), click.option("--all-starting", is_flag=True),
           click.option("--all-unusable", is_flag=True), click.option("--nodeid", multiple=True)])
def nodes_del(c, all_start_task_failed, all_starting, all_unusable, nodeid):
    run_action(c, fleet.action_pool_nodes_del, all_start_task_failed, all_starting, all_unusable, nodeid)


@_sub_cmd(pool_nodes, "reboot", "Reboot (re-prepare) nodes", [click.option("--all-start-task-failed", is_flag=True), click.option("--nodeid", multiple=True)])
def nodes_reboot(c, all_start_task_failed, nodeid):
    run_action(c, fleet.action_pool_nodes_reboot, all_start_task_failed, nodeid)


@_sub_cmd(pool_nodes, "ps", "List running task processes on all nodes")
def nodes_ps(c):
    run_action(c, fleet.action_pool_nodes_ps)


@_sub_cmd(pool_nodes, "zap", "Kill all task processes on all nodes", [click.option("--no-remove", is_flag=True), click.option("--stop", is_flag=True)])
def nodes_zap(c, no_remove, stop):
    run_action(c, fleet.action_pool_nodes_zap, no_remove, stop)


@_sub_cmd(pool_nodes, "prune", "Prune data of expired completed tasks", [click.option("--volumes", is_flag=True)])
def nodes_prune(c, volumes):
    run_action(c, fleet.action_pool_nodes_prune, volumes)


# ---------------------------------------------------------------------------------------------- jobs
@cli.group()
def jobs():
    """Jobs actions"""


@jobs.group("tasks")
def jobs_tasks():
    """Tasks actions"""


def _jobs_cmd(group, name, help_, opts=(), need=(ConfigType.Jobs,)):
    return _sub_cmd(group, name, help_, opts, need=need)


@_jobs_cmd(jobs, "add", "Add jobs",
           [click.option("--recreate", is_flag=True, help="Recreate existing jobs"), click.option("--tail", help="Tail a file of the last task"),
            click.option("--wait", is_flag=True, help="Run the node agent inline until all tasks finish"),
            click.option("--dry-run", is_flag=True, help="Validate and show the synthesised tasks without submitting")],
           need=(ConfigType.Jobs, ConfigType.Pool))
def jobs_add(c, recreate, tail, wait, dry_run):
    if not tail or c.raw:
        run_action(c, fleet.action_jobs_add, recreate, tail, wait, dry_run)
        return
    # --tail without --raw: stream the file's text as it is (the reference streams the task file to the terminal), then the summary
    try:
        out = fleet.action_jobs_add(c.ctx, recreate, tail, wait, dry_run)
    except (KeyError, ValueError, RuntimeError, OSError) as e:
        click.echo(f"ERROR: {e}", err=True)
        sys.exit(1)
    for jid, rec in out.items():
        if isinstance(rec, dict):
            text = rec.pop("tail", None)
            err = rec.pop("stderr_tail", None)
            if text is not None:
                click.echo(f"--- {jid}: {tail} ---")
                click.echo(text, nl=not str(text).endswith("\n"))
            if err:
                click.echo(f"--- {jid}: stderr (tail) ---", err=True)
                click.echo(err, err=True, nl=not str(err).endswith("\n"))
    emit(c, out)


@_jobs_cmd(jobs, "list", "List jobs", [click.option("--jobid"), click.option("--jobscheduleid")], need=())
def jobs_list(c, jobid, jobscheduleid):
    run_action(c, fleet.action_jobs_list, jobid, jobscheduleid)


_term_opts = [click.option("--all-jobs", is_flag=True), click.option("--all-jobschedules", is_flag=True), click.option("--jobid"),
              click.option("--jobscheduleid"), click.option("--termtasks", is_flag=True), click.option("--wait", is_flag=True)]


@_jobs_cmd(jobs, "term", "Terminate jobs", _term_opts, need=())
def jobs_term(c, all_jobs, all_jobschedules, jobid, jobscheduleid, termtasks, wait):
    run_action(c, fleet.action_jobs_term, all_jobs, all_jobschedules, jobid, jobscheduleid, termtasks, wait)


@_jobs_cmd(jobs, "del", "Delete jobs", _term_opts, need=())
def jobs_del(c, all_jobs, all_jobschedules, jobid, jobscheduleid, termtasks, wait):
    run_action(c, fleet.action_jobs_del, all_jobs, all_jobschedules, jobid, jobscheduleid, termtasks, wait)


@_jobs_cmd(jobs, "cmi", "Cleanup multi-instance jobs", [click.option("--delete", is_flag=True)])
def jobs_cmi(c, delete):
    run_action(c, fleet.action_jobs_cmi, delete)


_dis_opts = [click.option("--jobid"), click.option("--jobscheduleid"), click.option("--requeue", is_flag=True),
             click.option("--terminate", is_flag=True), click.option("--wait", is_flag=True)]


@_jobs_cmd(jobs, "migrate", "Migrate jobs to another pool", _dis_opts + [click.option("--poolid")], need=())
def jobs_migrate(c, jobid, jobscheduleid, requeue, terminate, wait, poolid):
    run_action(c, fleet.action_jobs_migrate, jobid, jobscheduleid, poolid, requeue, terminate, wait)


@_jobs_cmd(jobs, "disable", "Disable jobs", _dis_opts, need=())
def jobs_disable(c, jobid, jobscheduleid, requeue, terminate, wait):
    run_action(c, fleet.action_jobs_disable, jobid, jobscheduleid, requeue, terminate, wait)


@_jobs_cmd(jobs, "enable", "Enable jobs", [click.option("--jobid"), click.option("--jobscheduleid")], need=())
def jobs_enable(c, jobid, jobscheduleid):
    run_action(c, fleet.action_jobs_enable, jobid, jobscheduleid)


@_jobs_cmd(jobs, "stats", "Job statistics", [click.option("--jobid")], need=())
def jobs_stats(c, jobid):
    run_action(c, fleet.action_jobs_stats, jobid)


@_jobs_cmd(jobs_tasks, "list", "List tasks", [click.option("--all", "all_jobs", is_flag=True), click.option("--jobid"),
                                                click.option("--poll-until-tasks-complete", is_flag=True), click.option("--taskid")], need=())
def tasks_list(c, all_jobs, jobid, poll_until_tasks_complete, taskid):
    run_action(c, fleet.action_jobs_tasks_list, all_jobs, jobid, poll_until_tasks_complete, taskid)


@_jobs_cmd(jobs_tasks, "count", "Task counts", [click.option("--jobid")], need=())
def tasks_count(c, jobid):
    run_action(c, fleet.action_jobs_tasks_count, jobid)


@_jobs_cmd(jobs_tasks, "term", "Terminate tasks", [click.option("--force", is_flag=True), click.option("--jobid"), click.option("--taskid"),
                                                     click.option("--wait", is_flag=True)], need=())
def tasks_term(c, force, jobid, taskid, wait):
    run_action(c, fleet.action_jobs_tasks_term, force, jobid, taskid, wait)


@_jobs_cmd(jobs_tasks, "del", "Delete tasks", [click.option("--jobid"), click.option("--taskid"), click.option("--wait", is_flag=True)], need=())
def tasks_del(c, jobid, taskid, wait):
    run_action(c, fleet.action_jobs_tasks_del, jobid, taskid, wait)


# ---------------------------------------------------------------------------------------------- data
@cli.group()
def data():
    """Data actions"""


@data.group("files")
def data_files():
    """File actions"""


@_sub_cmd(data, "ingress", "Ingress data into the pool's shared volumes / storage", [click.option("--to-fs")], need=(),
          keep=BATCH + (ConfigType.RemoteFS,), extra_file_opts=(fs_option,))
def data_ingress(c, to_fs):
    run_action(c, fleet.action_data_ingress, to_fs)


@_sub_cmd(data_files, "list", "List files of tasks", [click.option("--jobid"), click.option("--taskid")], need=())
def files_list(c, jobid, taskid):
    run_action(c, fleet.action_data_files_list, jobid, taskid)


@_sub_cmd(data_files, "stream", "Stream a task file", [click.option("--disk", is_flag=True), click.option("--filespec", required=True)], need=())
def files_stream(c, disk, filespec):
    try:
        out = fleet.action_data_files_stream(c.ctx, disk, filespec)
    except (fleet.ActionError, fleet.BackendError) as e:
        click.echo(f"ERROR: {e}", err=True); sys.exit(1)
    if c.raw:
        emit(c, out)


@_sub_cmd(data_files, "task", "Retrieve task file(s)", [click.option("--all", is_flag=True), click.option("--filespec", required=True)], need=())
def files_task(c, all, filespec):
    run_action(c, fleet.action_data_files_task, all, filespec)


@_sub_cmd(data_files, "node", "Retrieve file(s) from a node", [click.option("--all", is_flag=True), click.option("--filespec", required=True)])
def files_node(c, all, filespec):
    run_action(c, fleet.action_data_files_node, all, filespec)


# ---------------------------------------------------------------------------------------------- diag / misc
@cli.group()
def diag():
    """Diagnostics actions"""


@diag.group("logs")
def diag_logs():
    """Diagnostic log actions"""


@_sub_cmd(diag_logs, "upload", "Collect node logs into the diagnostics container",
          [click.option("--cardinal", type=int), click.option("--generate-sas", is_flag=True), click.option("--nodeid"), click.option("--wait", is_flag=True)])
def diag_logs_upload(c, cardinal, generate_sas, nodeid, wait):
    run_action(c, fleet.action_diag_logs_upload, cardinal, generate_sas, nodeid, wait)


@cli.group()
def misc():
    """Miscellaneous actions"""


@_sub_cmd(misc, "tensorboard", "TensorBoard for a task's log directory", [click.option("--jobid"), click.option("--taskid"),
                                                                          click.option("--logdir"), click.option("--image")], need=())
def misc_tensorboard(c, jobid, taskid, logdir, image):
    run_action(c, fleet.action_misc_tensorboard, jobid, taskid, logdir, image)


@_sub_cmd(misc, "mirror-images", "Show which system/image artefacts are mirrored locally", need=())
def misc_mirror(c):
    run_action(c, fleet.action_misc_mirror_images)


# ---------------------------------------------------------------------------------------------- storage
@cli.group()
def storage():
    """Storage (local state) actions"""


@storage.group("sas")
def storage_sas():
    """SAS token actions"""


@_sub_cmd(storage, "clear", "Clear metadata of a pool", [click.option("--diagnostics-logs", is_flag=True), click.option("--poolid")], need=())
def storage_clear(c, diagnostics_logs, poolid):
    run_action(c, fleet.action_storage_clear, diagnostics_logs, poolid)


@_sub_cmd(storage, "del", "Delete metadata", [click.option("--clear-tables", is_flag=True), click.option("--diagnostics-logs", is_flag=True),
                                                click.option("--poolid")], need=())
def storage_del(c, clear_tables, diagnostics_logs, poolid):
    run_action(c, fleet.action_storage_del, clear_tables, diagnostics_logs, poolid)


@_sub_cmd(storage_sas, "create", "Create a SAS (a file:// URL locally)",
          [click.argument("storage_account"), click.argument("path"), click.option("--file", is_flag=True), click.option("--create", is_flag=True),
           click.option("--list", "list_", is_flag=True), click.option("--read", is_flag=True), click.option("--write", is_flag=True),
           click.option("--delete", is_flag=True)], need=())
def sas_create(c, storage_account, path, file, create, list_, read, write, delete):
    run_action(c, fleet.action_storage_sas_create, storage_account, path, file, create, list_, read, write, delete)


# ---------------------------------------------------------------------------------------------- keyvault / cert
@cli.group()
def keyvault():
    """KeyVault (local secret store) actions"""


@_sub_cmd(keyvault, "add", "Store the credentials config as a secret", [click.argument("name")], need=(ConfigType.Credentials,))
def keyvault_add(c, name):
    run_action(c, fleet.action_keyvault_add, name)


@_sub_cmd(keyvault, "del", "Delete a secret", [click.argument("name")], need=())
def keyvault_del(c, name):
    run_action(c, fleet.action_keyvault_del, name)


@_sub_cmd(keyvault, "list", "List secrets", need=())
def keyvault_list(c):
    run_action(c, fleet.action_keyvault_list)


@cli.group()
def cert():
    """Certificate actions"""


@_sub_cmd(cert, "create", "Create a certificate (PEM + PFX) for credential encryption", [click.option("--file-prefix"), click.option("--pfx-password")], need=())
def cert_create(c, file_prefix, pfx_password):
    run_action(c, fleet.action_cert_create, file_prefix, pfx_password)


@_sub_cmd(cert, "add", "Add a certificate to the local store", [click.option("--file"), click.option("--pem-no-certs", is_flag=True),
                                                                 click.option("--pem-public-key", is_flag=True), click.option("--pfx-password")], need=())
def cert_add(c, file, pem_no_certs, pem_public_key, pfx_password):
    run_action(c, fleet.action_cert_add, file, pem_no_certs, pem_public_key, pfx_password)


@_sub_cmd(cert, "list", "List certificates", need=())
def cert_list(c):
    run_action(c, fleet.action_cert_list)


@_sub_cmd(cert, "del", "Delete certificates", [click.option("--sha1", multiple=True)], need=())
def cert_del(c, sha1):
    run_action(c, fleet.action_cert_del, sha1)


# ---------------------------------------------------------------------------------------------- fs
@cli.group()
def fs():
    """Filesystem in the box (remote fs analogue) actions"""


@fs.group("cluster")
def fs_cluster():
    """Storage cluster actions"""


@fs.group("disks")
def fs_disks():
    """Managed disk actions"""


_FS_KEEP = (ConfigType.Credentials, ConfigType.Global, ConfigType.RemoteFS)


def _fs_cmd(group, name, help_, opts=(), need=(ConfigType.RemoteFS,)):
    return _sub_cmd(group, name, help_, opts, need=need, keep=_FS_KEEP, extra_file_opts=(fs_option,))


_sc_arg = click.argument("storage_cluster_id")


@_fs_cmd(fs_cluster, "add", "Create a storage cluster", [_sc_arg])
def fs_cluster_add(c, storage_cluster_id):
    run_action(c, fleet.action_fs_cluster_add, storage_cluster_id)


@_fs_cmd(fs_cluster, "orchestrate", "Create disks and the storage cluster", [_sc_arg])
def fs_cluster_orchestrate(c, storage_cluster_id):
    run_action(c, fleet.action_fs_cluster_orchestrate, storage_cluster_id)


@_fs_cmd(fs_cluster, "resize", "Resize a storage cluster", [_sc_arg])
def fs_cluster_resize(c, storage_cluster_id):
    run_action(c, fleet.action_fs_cluster_resize, storage_cluster_id)


@_fs_cmd(fs_cluster, "expand", "Expand a storage cluster with more disks", [_sc_arg, click.option("--no-rebalance", is_flag=True)])
def fs_cluster_expand(c, storage_cluster_id, no_rebalance):
    run_action(c, fleet.action_fs_cluster_expand, storage_cluster_id, no_rebalance)


@_fs_cmd(fs_cluster, "del", "Delete a storage cluster",
         [_sc_arg, click.option("--delete-resource-group", is_flag=True), click.option("--delete-data-disks", is_flag=True),
          click.option("--delete-virtual-network", is_flag=True), click.option("--generate-from-prefix", is_flag=True), click.option("--no-wait", is_flag=True)], need=())
def fs_cluster_del(c, storage_cluster_id, **kw):
    run_action(c, fleet.action_fs_cluster_del, storage_cluster_id, **kw)


@_fs_cmd(fs_cluster, "suspend", "Suspend a storage cluster", [_sc_arg, click.option("--no-wait", is_flag=True)], need=())
def fs_cluster_suspend(c, storage_cluster_id, no_wait):
    run_action(c, fleet.action_fs_cluster_suspend, storage_cluster_id, no_wait)


@_fs_cmd(fs_cluster, "start", "Start a suspended storage cluster", [_sc_arg, click.option("--no-wait", is_flag=True)], need=())
def fs_cluster_start(c, storage_cluster_id, no_wait):
    run_action(c, fleet.action_fs_cluster_start, storage_cluster_id, no_wait)


@_fs_cmd(fs_cluster, "status", "Storage cluster status", [_sc_arg, click.option("--detail", is_flag=True), click.option("--hosts", is_flag=True)], need=())
def fs_cluster_status(c, storage_cluster_id, detail, hosts):
    run_action(c, fleet.action_fs_cluster_status, storage_cluster_id, detail, hosts)


@_fs_cmd(fs_cluster, "ssh", "Run a command in the storage cluster directory",
         [_sc_arg, click.option("--cardinal", type=int), click.option("--hostname"), click.option("--tty", is_flag=True), click.argument("command", nargs=-1)], need=())
def fs_cluster_ssh(c, storage_cluster_id, cardinal, hostname, tty, command):
    run_action(c, fleet.action_fs_cluster_ssh, storage_cluster_id, cardinal, hostname, tty, command)


@_fs_cmd(fs_disks, "add", "Create managed disks (backing directories)")
def fs_disks_add(c):
    run_action(c, fleet.action_fs_disks_add)


@_fs_cmd(fs_disks, "del", "Delete managed disks", [click.option("--all", is_flag=True), click.option("--delete-resource-group", is_flag=True),
                                                   click.option("--name"), click.option("--resource-group"), click.option("--no-wait", is_flag=True),
                                                   click.option("--wait", is_flag=True, hidden=True)], need=())
def fs_disks_del(c, all, delete_resource_group, name, resource_group, no_wait, wait):
    run_action(c, fleet.action_fs_disks_del, all, delete_resource_group, name, resource_group, wait or not no_wait)


@_fs_cmd(fs_disks, "list", "List managed disks", [click.option("--resource-group"), click.option("--restrict-scope", is_flag=True)], need=())
def fs_disks_list(c, resource_group, restrict_scope):
    run_action(c, fleet.action_fs_disks_list, resource_group, restrict_scope)


# ---------------------------------------------------------------------------------------------- monitor
@cli.group()
def monitor():
    """Monitoring actions"""


_MON_KEEP = (ConfigType.Credentials, ConfigType.Global, ConfigType.Monitor, ConfigType.Pool)


def _mon_cmd(name, help_, opts=(), need=()):
    return _sub_cmd(monitor, name, help_, opts, need=need, keep=_MON_KEEP, extra_file_opts=(monitor_option,))


@_mon_cmd("create", "Create (start) the monitoring service", need=(ConfigType.Monitor,))
def monitor_create(c):
    run_action(c, fleet.action_monitor_create)


@_mon_cmd("add", "Add a resource to monitor", [click.option("--poolid", multiple=True), click.option("--remote-fs", multiple=True)])
def monitor_add(c, poolid, remote_fs):
    run_action(c, fleet.action_monitor_add, poolid, remote_fs)


@_mon_cmd("list", "List monitored resources")
def monitor_list(c):
    run_action(c, fleet.action_monitor_list)


@_mon_cmd("remove", "Remove a monitored resource", [click.option("--all", is_flag=True), click.option("--poolid", multiple=True), click.option("--remote-fs", multiple=True)])
def monitor_remove(c, all, poolid, remote_fs):
    run_action(c, fleet.action_monitor_remove, all, poolid, remote_fs)


@_mon_cmd("ssh", "Run a command in the monitoring service context", [click.option("--tty", is_flag=True), click.argument("command", nargs=-1)])
def monitor_ssh(c, tty, command):
    run_action(c, fleet.action_monitor_ssh, tty, command)


@_mon_cmd("suspend", "Stop the monitoring service", [click.option("--no-wait", is_flag=True)])
def monitor_suspend(c, no_wait):
    run_action(c, fleet.action_monitor_suspend, no_wait)


@_mon_cmd("start", "Start the monitoring service", [click.option("--no-wait", is_flag=True)])
def monitor_start(c, no_wait):
    run_action(c, fleet.action_monitor_start, no_wait)


@_mon_cmd("status", "Monitoring service status")
def monitor_status(c):
    run_action(c, fleet.action_monitor_status)


@_mon_cmd("destroy", "Destroy the monitoring service",
          [click.option("--delete-resource-group", is_flag=True), click.option("--delete-virtual-network", is_flag=True),
           click.option("--generate-from-prefix", is_flag=True), click.option("--no-wait", is_flag=True)])
def monitor_destroy(c, **kw):
    run_action(c, fleet.action_monitor_destroy, **kw)


# ---------------------------------------------------------------------------------------------- fed
@cli.group()
def fed():
    """Federation actions"""


@fed.group("proxy")
def fed_proxy():
    """Federation proxy (scheduler daemon) actions"""


@fed.group("pool")
def fed_pool():
    """Federation pool actions"""


@fed.group("jobs")
def fed_jobs():
    """Federation jobs actions"""


_FED_KEEP = BATCH + (ConfigType.Federation,)


def _fed_cmd(group, name, help_, opts=(), need=()):
    return _sub_cmd(group, name, help_, opts, need=need, keep=_FED_KEEP, extra_file_opts=(federation_option,))


_fid = click.argument("federation_id")


@_fed_cmd(fed_proxy, "create", "Start the federation scheduler daemon", need=(ConfigType.Federation,))
def fed_proxy_create(c):
    run_action(c, fleet.action_fed_proxy_create)


@_fed_cmd(fed_proxy, "ssh", "Federation proxy context", [click.option("--tty", is_flag=True), click.argument("command", nargs=-1)])
def fed_proxy_ssh(c, tty, command):
    run_action(c, fleet.action_fed_proxy_ssh, tty, command)


@_fed_cmd(fed_proxy, "suspend", "Stop the federation daemon", [click.option("--no-wait", is_flag=True)])
def fed_proxy_suspend(c, no_wait):
    run_action(c, fleet.action_fed_proxy_suspend, no_wait)


@_fed_cmd(fed_proxy, "start", "Start the federation daemon", [click.option("--no-wait", is_flag=True)])
def fed_proxy_start(c, no_wait):
    run_action(c, fleet.action_fed_proxy_start, no_wait)


@_fed_cmd(fed_proxy, "status", "Federation daemon status")
def fed_proxy_status(c):
    run_action(c, fleet.action_fed_proxy_status)


@_fed_cmd(fed_proxy, "destroy", "Destroy the federation daemon",
          [click.option("--delete-resource-group", is_flag=True), click.option("--delete-virtual-network", is_flag=True),
           click.option("--generate-from-prefix", is_flag=True), click.option("--no-wait", is_flag=True)])
def fed_proxy_destroy(c, **kw):
    run_action(c, fleet.action_fed_proxy_destroy, **kw)


@_fed_cmd(fed, "create", "Create a federation", [_fid, click.option("--force", is_flag=True), click.option("--no-unique-job-ids", is_flag=True)])
def fed_create(c, federation_id, force, no_unique_job_ids):
    run_action(c, fleet.action_fed_create, federation_id, force, no_unique_job_ids)


@_fed_cmd(fed, "list", "List federations", [click.option("--federation-id", multiple=True)])
def fed_list(c, federation_id):
    run_action(c, fleet.action_fed_list, federation_id)


@_fed_cmd(fed, "destroy", "Destroy a federation", [_fid])
def fed_destroy(c, federation_id):
    run_action(c, fleet.action_fed_destroy, federation_id)


@_fed_cmd(fed_pool, "add", "Add pool(s) to a federation", [_fid, click.option("--batch-service-url"), click.option("--pool-id", multiple=True)])
def fed_pool_add(c, federation_id, batch_service_url, pool_id):
    run_action(c, fleet.action_fed_pool_add, federation_id, batch_service_url, pool_id)


@_fed_cmd(fed_pool, "remove", "Remove pool(s) from a federation",
          [_fid, click.option("--all", is_flag=True), click.option("--batch-service-url"), click.option("--pool-id", multiple=True)])
def fed_pool_remove(c, federation_id, all, batch_service_url, pool_id):
    run_action(c, fleet.action_fed_pool_remove, federation_id, all, batch_service_url, pool_id)


@_fed_cmd(fed_jobs, "add", "Submit jobs to a federation", [_fid], need=(ConfigType.Jobs, ConfigType.Pool))
def fed_jobs_add(c, federation_id):
    run_action(c, fleet.action_fed_jobs_add, federation_id)


@_fed_cmd(fed_jobs, "list", "List jobs / queued / blocked actions of a federation",
          [_fid, click.option("--blocked", is_flag=True), click.option("--job-id"), click.option("--jobschedule-id"), click.option("--queued", is_flag=True)])
def fed_jobs_list(c, federation_id, blocked, job_id, jobschedule_id, queued):
    run_action(c, fleet.action_fed_jobs_list, federation_id, blocked, job_id, jobschedule_id, queued)


_fj = [_fid, click.option("--all-jobs", is_flag=True), click.option("--all-jobschedules", is_flag=True),
       click.option("--job-id", multiple=True), click.option("--jobschedule-id", "--job-schedule-id", "job_schedule_id", multiple=True)]


@_fed_cmd(fed_jobs, "term", "Terminate federation jobs", _fj + [click.option("--force", is_flag=True)])
def fed_jobs_term(c, federation_id, all_jobs, all_jobschedules, job_id, job_schedule_id, force):
    run_action(c, fleet.action_fed_jobs_term, federation_id, all_jobs, all_jobschedules, force, job_id, job_schedule_id)


@_fed_cmd(fed_jobs, "del", "Delete federation jobs", _fj)
def fed_jobs_del(c, federation_id, all_jobs, all_jobschedules, job_id, job_schedule_id):
    run_action(c, fleet.action_fed_jobs_del, federation_id, all_jobs, all_jobschedules, job_id, job_schedule_id)


@_fed_cmd(fed_jobs, "zap", "Remove a queued/blocked action by unique id", [_fid, click.option("--unique-id", required=True)])
def fed_jobs_zap(c, federation_id, unique_id):
    run_action(c, fleet.action_fed_jobs_zap, federation_id, unique_id)


# ---------------------------------------------------------------------------------------------- slurm
@cli.group()
def slurm():
    """Slurm on the box actions"""


@slurm.group("ssh")
def slurm_ssh():
    """Slurm shell actions"""


@slurm.group("cluster")
def slurm_cluster():
    """Slurm cluster actions"""


_SL_KEEP = (ConfigType.Credentials, ConfigType.Global, ConfigType.Slurm, ConfigType.RemoteFS, ConfigType.Pool)


def _sl_cmd(group, name, help_, opts=(), need=(ConfigType.Slurm,)):
    return _sub_cmd(group, name, help_, opts, need=need, keep=_SL_KEEP, extra_file_opts=(slurm_option, fs_option))


_ssh_opts = [click.option("--offset", type=int), click.option("--tty", is_flag=True), click.argument("command", nargs=-1)]


@_sl_cmd(slurm_ssh, "controller", "Controller context", _ssh_opts)
def slurm_ssh_controller(c, offset, tty, command):
    run_action(c, fleet.action_slurm_ssh, "controller", offset, None, tty, command)


@_sl_cmd(slurm_ssh, "login", "Login node context", _ssh_opts)
def slurm_ssh_login(c, offset, tty, command):
    run_action(c, fleet.action_slurm_ssh, "login", offset, None, tty, command)


@_sl_cmd(slurm_ssh, "node", "Compute node context", [click.option("--node-name", required=True), click.option("--tty", is_flag=True), click.argument("command", nargs=-1)])
def slurm_ssh_node(c, node_name, tty, command):
    run_action(c, fleet.action_slurm_ssh, "node", None, node_name, tty, command)


@_sl_cmd(slurm_cluster, "create", "Create the slurm <-> GPU partition mapping")
def slurm_cluster_create(c):
    run_action(c, fleet.action_slurm_cluster_create)


@_sl_cmd(slurm_cluster, "orchestrate", "Create fs cluster (optional) and slurm mapping", [click.option("--storage-cluster-id")])
def slurm_cluster_orchestrate(c, storage_cluster_id):
    run_action(c, fleet.action_slurm_cluster_orchestrate, storage_cluster_id)


_sl_state = [click.option("--no-controller-nodes", is_flag=True), click.option("--no-login-nodes", is_flag=True), click.option("--no-wait", is_flag=True)]


@_sl_cmd(slurm_cluster, "suspend", "Suspend the slurm cluster", _sl_state)
def slurm_cluster_suspend(c, no_controller_nodes, no_login_nodes, no_wait):
    run_action(c, fleet.action_slurm_cluster_suspend, no_controller_nodes, no_login_nodes, no_wait)


@_sl_cmd(slurm_cluster, "start", "Start the slurm cluster", _sl_state)
def slurm_cluster_start(c, no_controller_nodes, no_login_nodes, no_wait):
    run_action(c, fleet.action_slurm_cluster_start, no_controller_nodes, no_login_nodes, no_wait)


@_sl_cmd(slurm_cluster, "status", "Slurm cluster status")
def slurm_cluster_status(c):
    run_action(c, fleet.action_slurm_cluster_status)


@_sl_cmd(slurm_cluster, "destroy", "Destroy the slurm cluster mapping",
         [click.option("--delete-resource-group", is_flag=True), click.option("--delete-virtual-network", is_flag=True),
          click.option("--generate-from-prefix", is_flag=True), click.option("--no-wait", is_flag=True)])
def slurm_cluster_destroy(c, **kw):
    run_action(c, fleet.action_slurm_cluster_destroy, **kw)


def leaf_commands(group=None, prefix=()) -> list:
    """Every leaf command path (used by the docs and the parity test)."""
    group = group or cli
    out = []
    for name, cmd in sorted(group.commands.items()):
        if isinstance(cmd, click.Group):
            out += leaf_commands(cmd, prefix + (name,))
        else:
            out.append(" ".join(prefix + (name,)))
    return out


def main(argv=None):
    cli(args=argv, prog_name="shipyard", obj=CliContext())


if __name__ == "__main__":
    main()
