* a small mixed-integer problem touching every section and bound type
NAME          SMALLMIX
ROWS
 N  COST
 L  LIM1
 G  LIM2
 E  MYEQN
 G  RNG1
 L  RNG2
 E  RNG3
 E  RNG4
 N  SECONDOBJ
COLUMNS
    X1        COST         1.0   LIM1         1.0
    X1        LIM2         1.0
    MARKER    'MARKER'     'INTORG'
    Y1        COST         2.0   LIM1         1.0
    Y1        MYEQN       -1.0
    Y2        RNG1         3.0   RNG2         1.5
    MARKER    'MARKER'     'INTEND'
    X2        COST        -1.0   MYEQN        1.0
    X2        RNG3         2.0   RNG4        -2.0
    X3        RNG1         1.0   RNG2         1.0
    X3        RNG3         1.0   RNG4         1.0
    X3        SECONDOBJ    5.0
    X4        LIM2         4.0
    X4        LIM2        -4.0
    X5        LIM1         0.5   LIM1         0.25
    X6        RNG4         1e-10
    B1        LIM2         1.0   COST         0.125
RHS
    RHS1      COST        -7.5
    RHS1      LIM1         4.0   LIM2         1.0
    RHS1      MYEQN        7.0
    RHS1      RNG1         2.0   RNG2        10.0
    RHS1      RNG3         3.0   RNG4         5.0
    RHS2      LIM1        99.0
RANGES
    RNG       RNG1         4.0   RNG2        -6.0
    RNG       RNG3         2.5
    RNG       RNG4        -1.5
    OTHER     LIM1         1.0
BOUNDS
 UP BND       X1           4.0
 LO BND       Y1          -1.0
 UP BND       Y1           1.0
 UP BND       X2          -2.0
 FR BND       X3
 MI BND       X4
 PL BND       X4
 FX BND       X5           2.5
 BV BND       B1
 LI BND       Y2           1
 UI BND       Y2           6
 UP OTHERBND  X1          44.0
 LO BND       X6           0.5
 UP BND       X6          -0.5
ENDATA
