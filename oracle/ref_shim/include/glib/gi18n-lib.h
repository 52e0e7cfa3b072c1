/* shim: gettext markers only */
#ifndef SHIM_GI18N_H
#define SHIM_GI18N_H
#define _(S) (S)
#define N_(S) (S)
#endif
