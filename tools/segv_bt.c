/* LD_PRELOAD helper for GPU-box debugging: prints a native backtrace of the faulting thread on SIGSEGV / SIGABRT / SIGBUS.
 * gcc -shared -fPIC -o build/segv_bt.so tools/segv_bt.c ; run python with -p no:faulthandler so that the handler stays. */
#define _GNU_SOURCE
#include <execinfo.h>
#include <fcntl.h>
#include <stdlib.h>
#include <signal.h>
#include <stdio.h>
#include <string.h>
#include <unistd.h>

static int out_fd = -1;

static void on_fault(int sig, siginfo_t *si, void *uc)
{
    void *bt[96];
    char line[128];
    (void)uc;
    int n = snprintf(line, sizeof line, "\n[segv_bt] signal %d at address %p\n", sig, si ? si->si_addr : (void *)0);
    if (write(2, line, (size_t)n) < 0) _exit(140);
    n = backtrace(bt, 96);
    backtrace_symbols_fd(bt, n, 2);
    if (out_fd >= 0) {   /* pytest captures fd 2: a file of our own as well */
        backtrace_symbols_fd(bt, n, out_fd);
        close(out_fd);
    }
    _exit(139);
}

__attribute__((constructor)) static void install(void)
{
    struct sigaction sa;
    const char *f = getenv("SEGV_BT_FILE");
    if (f) out_fd = open(f, O_WRONLY | O_CREAT | O_APPEND, 0644);
    memset(&sa, 0, sizeof sa);
    sa.sa_sigaction = on_fault;
    sa.sa_flags = SA_SIGINFO;
    sigaction(SIGSEGV, &sa, 0);
    sigaction(SIGBUS, &sa, 0);
    sigaction(SIGABRT, &sa, 0);
}
