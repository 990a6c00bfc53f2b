#pragma once
typedef int HYPRE_Int; typedef int HYPRE_BigInt; typedef double HYPRE_Real; typedef double HYPRE_Complex;
